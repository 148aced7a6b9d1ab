"""One conv layer launched a few times (for an ncu capture of a single kernel): python tools/prof_one.py [bf16|fp32] [shape index]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

from yolopoint_b200._lib import YP_ALGO_TCGEN05, YP_FMT_BF16, YP_FMT_F32X2  # noqa: E402
import test_gpu_conv as T  # noqa: E402
from conv_occupancy import SHAPES  # noqa: E402

fmt = YP_FMT_BF16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else YP_FMT_F32X2
c = SHAPES[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
for _ in range(3):
    try:
        T.run_case(c, fmt, YP_ALGO_TCGEN05)
    except AssertionError as e:
        print(e)
torch.cuda.synchronize()
print("done", c)
