"""Host-side logic of the training-step convolutions (yolopoint_b200/train.py) on the CPU: the operand layouts handed to
yp_conv2d_nhwc_fwd -- packed forward weights, transposed / tap-flipped data-gradient weights, the four parity classes of a stride-2
data gradient with their custom tap lists and single-parity stores -- are executed by a torch interpreter of the C-ABI contract
(include/yolopoint_b200.h: weight [cout][tap*Cin + cin], out[p] = sum_t W[:, t, :] . in[p*s + offset_t], YP_UP_PARITY stores) and
compared with autograd.  No GPU, no library call."""
import pytest
import torch
import torch.nn.functional as F

from yolopoint_b200 import train as T
from yolopoint_b200._lib import YP_UP_PARITY


def _interp_launch(x, wp, y, ksize, stride, n_taps=0, dh=(), dw=(), up=1, out_hw=None):
    """torch statement of yp_conv2d_nhwc_fwd for bf16 views (bias / activation off), writing into y like the kernel does."""
    B, Ci, H, W = x.shape
    co = wp.shape[0]
    if ksize == 1:
        taps = [(0, 0)]
    elif ksize == 3:
        taps = [(kh - 1, kw - 1) for kh in range(3) for kw in range(3)]
    else:
        taps = list(zip(dh[:n_taps], dw[:n_taps]))
    Ho, Wo = (H // stride, W // stride)
    xf = F.pad(x.float(), (8, 8, 8, 8))
    wf = wp.float().view(co, len(taps), Ci)
    out = torch.zeros(B, co, Ho, Wo)
    for t, (oh, ow) in enumerate(taps):
        sl = xf[:, :, 8 + oh: 8 + oh + H: stride, 8 + ow: 8 + ow + W: stride][:, :, :Ho, :Wo]
        out += torch.einsum("bchw,oc->bohw", sl, wf[:, t])
    out = out.to(torch.bfloat16)
    if up == 1:
        y.copy_(out)
    else:
        pp = up - YP_UP_PARITY
        y[:, :, (pp >> 1)::2, (pp & 1)::2] = out


@pytest.fixture
def cpu_kernels(monkeypatch):
    monkeypatch.setattr(T, "_launch_conv", _interp_launch)


@pytest.mark.parametrize("Ci,Co,H,W,k,s", [(16, 32, 10, 12, 1, 1), (16, 16, 9, 14, 3, 1), (32, 16, 12, 16, 3, 2), (16, 48, 8, 8, 3, 2)])
def test_forward_and_data_gradient_layouts(cpu_kernels, Ci, Co, H, W, k, s):
    g = torch.Generator().manual_seed(Ci + Co + k)
    x = torch.randn(2, Ci, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).to(torch.bfloat16).float()
    dy = torch.randn(2, Co, H // s, W // s, generator=g).to(torch.bfloat16)
    xr, wr = x.float().requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, s, k // 2)
    yr.backward(dy.float())
    y = T.conv_forward(T._cl(x), w, s)
    dx = T.conv_dgrad(T._cl(dy), w, s, H, W)
    assert float((y.float() - yr.detach()).abs().max()) < 2e-2 * float(yr.abs().max())
    assert float((dx.float() - xr.grad).abs().max()) < 2e-2 * float(xr.grad.abs().max())
    # the pre-packed operands (WeightPack) are the same matrices
    pk = T.WeightPack(lambda: w, s, need_dgrad=True)
    pk.refresh()
    dx2 = T.conv_dgrad(T._cl(dy), w, s, H, W, pk)
    assert torch.equal(dx2, dx) and torch.equal(T.conv_forward(T._cl(x), w, s, pk), y)
    w2 = w * 0.5
    pk.make_weight = lambda: w2
    pk.refresh()          # in place: same buffers, new values
    assert float((T.conv_forward(T._cl(x), w2, s, pk).float() - 0.5 * yr.detach()).abs().max()) < 2e-2 * float(yr.abs().max())


def test_stem_space_to_depth_is_the_6x6_stride2_conv(cpu_kernels):
    """6x6 s2 p2 on 3 channels == 3x3 s1 p1 on the 2x2 space-to-depth image with the remapped weight (both rearrangements are torch ops)."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 16, 24, generator=g).to(torch.bfloat16).float()
    w = (torch.randn(8, 3, 6, 6, generator=g) / 10).to(torch.bfloat16).float()
    xs, ws = T._stem_s2d(x, w)
    ref = F.conv2d(x, w, None, 2, 2)
    got = F.conv2d(xs, ws, None, 1, 1)
    assert torch.allclose(got, ref, atol=1e-5)
    assert torch.equal(T._stem_weight(w), ws)
