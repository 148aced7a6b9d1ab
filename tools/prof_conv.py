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
    dict(B=4, H=92, W=160, Cin=192, Cout=384, k=3, s=2),              # YOLOPoint-M batch 4: Conv4 (best multi-wave layer)
    dict(B=8, H=80, W=80, Cin=128, Cout=128, k=3, s=1, act=False, nobias=True),   # YOLOPoint-L training, batch 8 (15 per pass)
    dict(B=8, H=80, W=80, Cin=256, Cout=256, k=3, s=1, act=False, nobias=True),   # YOLOPoint-L training: ConvDesc
]
if __name__ == "__main__":
    fmts = [YP_FMT_F32X2, YP_FMT_BF16]
    for fmt in fmts:
        for c in SHAPES:
            for _ in range(2):
                T.run_case(dict(c), fmt, YP_ALGO_TCGEN05)
    # weight gradient of the two training layers (wgrad_tc_kernel)
    import torch
    from yolopoint_b200 import train as TR
    for ci, co in ((128, 128), (256, 256)):
        x = torch.randn(8, ci, 80, 80, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dy = torch.randn(8, co, 80, 80, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        for _ in range(2):
            TR.conv_wgrad(x, dy, 3, 1)
    torch.cuda.synchronize()
    print("done")
