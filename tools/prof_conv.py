"""Launch a few representative conv layers (for `ncu --set full -k regex:conv_tc`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_conv as T  # noqa: E402
from yolopoint_b200._lib import YP_ALGO_TCGEN05, YP_FMT_BF16, YP_FMT_F32X2  # noqa: E402

SHAPES = [
    dict(B=1, H=80, W=80, Cin=64, Cout=64, k=3, s=1, res=True),       # Bottleneck 3x3 at stride 8 (5 per frame)
    dict(B=1, H=80, W=80, Cin=128, Cout=128, k=1, s=1),               # C3 cv1||cv2 / cv3 at stride 8
    dict(B=1, H=20, W=20, Cin=256, Cout=256, k=3, s=1, res=True),     # deep 3x3 at stride 32 (split-K)
    dict(B=4, H=184, W=320, Cin=48, Cout=96, k=3, s=2),               # YOLOPoint-M 1280x736 batch 4: Conv2
]
if __name__ == "__main__":
    fmts = [YP_FMT_F32X2, YP_FMT_BF16]
    for fmt in fmts:
        for c in SHAPES:
            for _ in range(2):
                T.run_case(dict(c), fmt, YP_ALGO_TCGEN05)
    print("done")
