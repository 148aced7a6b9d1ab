"""A few training-step kernels for an ncu capture: persistent conv (forward / data gradient), weight gradient, BN + SiLU passes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from yolopoint_b200 import train as T  # noqa: E402

CL = torch.channels_last
for ci, co, hw, k, s in ((128, 128, 80, 3, 1), (256, 256, 80, 3, 1), (64, 128, 320, 3, 2), (256, 256, 80, 1, 1)):
    x = torch.randn(8, ci, hw, hw, device="cuda").to(torch.bfloat16).contiguous(memory_format=CL)
    w = torch.randn(co, ci, k, k, device="cuda") / (ci * k * k) ** 0.5
    y = T.conv_forward(x, w, s)
    dy = torch.randn_like(y)
    for _ in range(2):
        T.conv_forward(x, w, s)
        T.conv_dgrad(dy, w, s, hw, hw)
        T.conv_wgrad(x, dy, k, s)
bn = torch.nn.BatchNorm2d(128, eps=1e-3, momentum=0.03).cuda()
y = torch.randn(8, 128, 80, 80, device="cuda").to(torch.bfloat16).contiguous(memory_format=CL).requires_grad_(True)
for _ in range(2):
    out = T.bn_act_tc(y, bn, True)
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
print("done")
