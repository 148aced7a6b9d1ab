"""GPU parity of yp_conv2d_nhwc_fwd (tcgen05 path and the SIMT cross-check) through the C ABI.

Reference for a floating-point kernel: plain PyTorch fp32 conv2d on the same device with TF32 disabled,
fed the exact operand values the kernel sees (hi+lo planes / bf16-rounded values)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from yolopoint_b200 import _lib
from yolopoint_b200._lib import (YP_ACT_NONE, YP_ACT_SILU, YP_ALGO_SIMT, YP_ALGO_TCGEN05, YP_EPI_L2NORM, YP_FMT_BF16, YP_FMT_F32,
                                  YP_FMT_F32X2, YpConvDesc)
from yolopoint_b200.engine import make_view, split_tf32

pytestmark = pytest.mark.gpu


def _mk_act(vals: torch.Tensor, fmt: int, ctot: int, c_off: int):
    """vals [B,H,W,C] fp32 (cuda) -> buffer [P,B,H,W,ctot] holding vals at channel offset c_off, rest = 7.0 sentinel."""
    B, H, W, Cc = vals.shape
    if fmt == YP_FMT_BF16:
        buf = torch.full((1, B, H, W, ctot), 7.0, dtype=torch.bfloat16, device=vals.device)
        buf[0, ..., c_off:c_off + Cc] = vals.to(torch.bfloat16)
        eff = buf[0, ..., c_off:c_off + Cc].float()
    else:
        buf = torch.full((2, B, H, W, ctot), 7.0, dtype=torch.float32, device=vals.device)
        buf[:, ..., c_off:c_off + Cc] = split_tf32(vals)
        eff = buf[0, ..., c_off:c_off + Cc] + buf[1, ..., c_off:c_off + Cc]
    return buf, eff


def _read(buf: torch.Tensor, fmt: int, c_off: int, Cc: int):
    t = buf[..., c_off:c_off + Cc].float()
    return t[0] + t[1] if fmt == YP_FMT_F32X2 else t[0]


CASES = [
    # B, H, W, Cin, Cout, k, s, act, res, l2, up, out_fmt_plain
    dict(B=1, H=16, W=16, Cin=32, Cout=32, k=1, s=1),
    dict(B=2, H=20, W=20, Cin=64, Cout=64, k=3, s=1, res=True),
    dict(B=1, H=40, W=24, Cin=32, Cout=64, k=3, s=2),
    dict(B=1, H=80, W=80, Cin=128, Cout=128, k=3, s=1, act=False, l2=True, plain=True),
    dict(B=1, H=20, W=20, Cin=256, Cout=128, k=1, s=1, up=True),
    dict(B=1, H=24, W=40, Cin=16, Cout=16, k=3, s=1),
    dict(B=2, H=12, W=20, Cin=48, Cout=96, k=3, s=2),
    dict(B=1, H=8, W=8, Cin=512, Cout=256, k=1, s=1, act=False, plain=True),
    dict(B=1, H=80, W=80, Cin=128, Cout=80, k=1, s=1, act=False, plain=True, nobias=True),
    dict(B=3, H=34, W=18, Cin=64, Cout=32, k=3, s=1, res=True),
    dict(B=1, H=20, W=20, Cin=256, Cout=256, k=3, s=1, res=True),     # deep layer on a small map: split-K
    dict(B=1, H=40, W=40, Cin=128, Cout=128, k=3, s=2),
    dict(B=1, H=20, W=20, Cin=1024, Cout=512, k=1, s=1, split=4),
    dict(B=1, H=80, W=80, Cin=128, Cout=128, k=3, s=1, act=False, l2=True, plain=True, nobias=True, split=3),   # L2-norm head with split-K
    dict(B=1, H=40, W=40, Cin=128, Cout=128, k=3, s=1, res=True, split=4, tile=64),
    # several waves of tiles: in bf16 these run on the persistent kernel (double-buffered TMEM accumulators)
    dict(B=8, H=80, W=80, Cin=128, Cout=128, k=3, s=1, res=True),                  # patch mode, 432 tiles
    dict(B=8, H=80, W=80, Cin=128, Cout=256, k=1, s=1),                            # 400 x 1 or 2 N tiles
    dict(B=6, H=96, W=160, Cin=64, Cout=128, k=3, s=2, act=False, nobias=True),    # stride 2, 360 tiles
    dict(B=5, H=72, W=72, Cin=64, Cout=80, k=1, s=1, act=False, plain=True),       # fp32 output, odd tile counts
    dict(B=8, H=80, W=80, Cin=64, Cout=64, k=3, s=1, up=True),                     # two destinations (+ 2x upsample)
]


def run_case(c, fmt, algo, shared_ws=None):
    L = _lib.lib(require_device=True)
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(1234)
    B, H, W, Cin, Cout, k, s = c["B"], c["H"], c["W"], c["Cin"], c["Cout"], c["k"], c["s"]
    act = YP_ACT_SILU if c.get("act", True) else YP_ACT_NONE
    Ho, Wo = H // s, W // s
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    bias = None if c.get("nobias") else torch.randn(Cout, generator=g).to(dev)
    in_buf, x_eff = _mk_act(x, fmt, Cin + 16, 16)                   # input is a channel slice of a wider buffer
    wk = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous()
    if fmt == YP_FMT_BF16:
        wp = wk.to(torch.bfloat16).unsqueeze(0).contiguous()
        w_eff = wp[0].float()
    else:
        wp = split_tf32(wk).contiguous()
        w_eff = wp[0] + wp[1]
    out_fmt = YP_FMT_F32 if c.get("plain") else fmt
    planes = 2 if out_fmt == YP_FMT_F32X2 else 1
    odt = torch.bfloat16 if out_fmt == YP_FMT_BF16 else torch.float32
    up = 2 if c.get("up") else 1
    out_buf = torch.full((planes, B, Ho * up, Wo * up, Cout + 32), 5.0, dtype=odt, device=dev)   # slice at channel offset 32
    out2_buf = torch.full((planes, B, Ho, Wo, Cout), 5.0, dtype=odt, device=dev) if c.get("up") else None
    res_buf = res_eff = None
    if c.get("res"):
        r = torch.randn(B, Ho, Wo, Cout, generator=g).to(dev)
        res_buf, res_eff = _mk_act(r, fmt, Cout, 0)
    d = YpConvDesc()
    d.in_ = make_view(in_buf, fmt, 16, Cin)
    d.weight, d.bias = wp.data_ptr(), (bias.data_ptr() if bias is not None else None)
    d.ksize, d.stride, d.cout, d.act = k, s, Cout, act
    d.epilogue = YP_EPI_L2NORM if c.get("l2") else 0
    if res_buf is not None:
        d.residual = make_view(res_buf, fmt, 0, Cout)
    d.n_out = 2 if c.get("up") else 1
    d.out[0] = make_view(out_buf, out_fmt, 32, Cout, upsample=up)
    if c.get("up"):
        d.out[1] = make_view(out2_buf, out_fmt, 0, Cout)
    d.algo = algo
    d.split_k = c.get("split", 0)
    d.tile_n = c.get("tile", 0)
    ws = None
    if algo == YP_ALGO_TCGEN05:
        nbytes = int(L.yp_conv2d_workspace_bytes(C.byref(d)))
        if nbytes and shared_ws is not None:
            assert nbytes <= shared_ws.numel()
            d.workspace, d.workspace_bytes = shared_ws.data_ptr(), shared_ws.numel()
        elif nbytes:
            ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            d.workspace, d.workspace_bytes = ws.data_ptr(), nbytes
    for _ in range(2):   # twice: the split-K arrival counters must reset themselves
      _lib.check(L.yp_conv2d_nhwc_fwd(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    # reference
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    wr = w_eff.view(Cout, k, k, Cin).permute(0, 3, 1, 2).contiguous()
    y = F.conv2d(x_eff.permute(0, 3, 1, 2).double(), wr.double(), None if bias is None else bias.double(), stride=s, padding=k // 2)
    if act:
        y = F.silu(y)
    y = y.permute(0, 2, 3, 1)
    if res_eff is not None:
        y = y + res_eff.double()
    if c.get("l2"):
        y = y / y.norm(dim=-1, keepdim=True)
    y = y.float()
    got = _read(out_buf, out_fmt, 32, Cout)
    if up == 2:
        y_up = y.repeat_interleave(2, 1).repeat_interleave(2, 2)
        got2 = _read(out2_buf, out_fmt, 0, Cout)
    else:
        y_up, got2 = y, None
    tol = 2e-2 if fmt == YP_FMT_BF16 else 2e-5
    scale = float(y.abs().max()) + 1e-6
    err = float((got - y_up).abs().max()) / scale
    assert err < tol, f"{c} fmt={fmt} algo={algo}: rel err {err:.3e}"
    if got2 is not None:
        err2 = float((got2 - y).abs().max()) / scale
        assert err2 < tol, f"{c} second output rel err {err2:.3e}"
    # untouched channels of the destination buffer keep their sentinel
    assert float(out_buf[..., :32].float().min()) == 5.0 and float(out_buf[..., :32].float().max()) == 5.0
    return err


@pytest.mark.parametrize("ci", range(len(CASES)))
@pytest.mark.parametrize("fmt", [YP_FMT_F32X2, YP_FMT_BF16])
def test_conv_tcgen05(ci, fmt):
    run_case(CASES[ci], fmt, YP_ALGO_TCGEN05)


@pytest.mark.parametrize("ci", [1, 2, 3, 4])
def test_conv_simt_crosscheck(ci):
    run_case(CASES[ci], YP_FMT_F32X2, YP_ALGO_SIMT)


def test_split_k_layers_share_one_workspace():
    """Layers of one lane share a split-K workspace (engine.ShapePlan): the arrival counters of a layer with many tiles
    must not alias the partial sums an earlier layer with few tiles left behind (regression: per-layer counter area)."""
    ws = torch.zeros(192 << 20, dtype=torch.uint8, device="cuda")
    few = dict(B=1, H=20, W=20, Cin=1024, Cout=512, k=1, s=1, split=4, tile=128)       # 4 x 4 tiles
    many = dict(B=1, H=40, W=40, Cin=512, Cout=256, k=1, s=1, split=2, tile=32)        # 13 x 8 tiles
    more = dict(B=2, H=48, W=80, Cin=384, Cout=384, k=1, s=1, split=12, tile=64)       # 60 x 6 tiles
    for c in (few, many, few, more, many):
        run_case(c, YP_FMT_F32X2, YP_ALGO_TCGEN05, shared_ws=ws)


# Throughput plan (tile_n = YP_TILE_WIDE = -1): 3xTF32 layers with 128-byte store chunks run on conv_tc_drain_kernel (persistent,
# accumulators drained into registers every 16 MMAs); the others take the chain-bounded wide plans of conv_tc_kernel.
WIDE_CASES = [
    dict(B=1, H=16, W=16, Cin=32, Cout=32, k=1, s=1),                              # one tile, one round
    dict(B=1, H=80, W=80, Cin=128, Cout=128, k=1, s=1),                            # Nt = 128, 50 tiles
    dict(B=2, H=20, W=20, Cin=64, Cout=64, k=3, s=1, res=True),                    # patch mode + residual
    dict(B=1, H=20, W=20, Cin=256, Cout=256, k=3, s=1, res=True),                  # deep K (288 k-steps = 18 rounds), two N tiles
    dict(B=1, H=40, W=40, Cin=128, Cout=128, k=3, s=2),                            # stride 2 (parity views)
    dict(B=1, H=20, W=20, Cin=1024, Cout=512, k=1, s=1),                           # 32 k-blocks, four N tiles
    dict(B=1, H=20, W=20, Cin=256, Cout=128, k=1, s=1, up=True),                   # two destinations (+ 2x upsample)
    dict(B=1, H=8, W=8, Cin=512, Cout=256, k=1, s=1, act=False, plain=True),       # fp32 output
    dict(B=8, H=80, W=80, Cin=128, Cout=128, k=3, s=1, res=True),                  # 400+ tiles: several tiles per persistent CTA
    dict(B=6, H=96, W=160, Cin=64, Cout=96, k=3, s=2, act=False, nobias=True),     # Nt = 96, 360 tiles
    dict(B=1, H=320, W=320, Cin=16, Cout=32, k=3, s=1),                            # the stem's geometry: 64-byte K rows, 856 tiles
    dict(B=1, H=80, W=80, Cin=128, Cout=80, k=1, s=1, act=False, plain=True, nobias=True),   # 64-byte store chunks: not drained (conv_tc_kernel)
    dict(B=1, H=80, W=80, Cin=128, Cout=128, k=3, s=1, act=False, l2=True, plain=True),      # L2-norm head: not drained
]


@pytest.mark.parametrize("ci", range(len(WIDE_CASES)))
def test_conv_wide_plan(ci):
    c = dict(WIDE_CASES[ci], tile=-1)
    err = run_case(c, YP_FMT_F32X2, YP_ALGO_TCGEN05)
    print(f"wide plan case {ci}: rel err {err:.2e}")


@pytest.mark.parametrize("ci", [1, 3, 8, 9, 10])
def test_conv_wide_plan_shared_grid(ci):
    """tile_n = -3: the same plan with the persistent grid capped at a third of the SMs (several tiles per CTA: what bench.py's
    headline configuration launches)."""
    c = dict(WIDE_CASES[ci], tile=-3)
    err = run_case(c, YP_FMT_F32X2, YP_ALGO_TCGEN05)
    print(f"wide plan (grid / 3) case {ci}: rel err {err:.2e}")
