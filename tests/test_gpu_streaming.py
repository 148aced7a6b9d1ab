"""GPU parity of the streaming (HBM-bound) kernels' fast paths: vectorised NHWC heatmap, division-free Detect decode, two-pixel
frame conversion.  Each is compared bit-exactly with the kernel's general path or with the oracle restatement
(src/models/yolo.py:49-81, src/utils/utils.py:232-262, src/demo.py:130-132 of the reference)."""
import ctypes
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import yolopoint_oracle as O
from yolopoint_b200 import _lib, ops
from yolopoint_b200._lib import YP_FMT_BF16, YP_FMT_F32X2
from yolopoint_b200.engine import make_view, split_tf32

pytestmark = pytest.mark.gpu


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("B,Hc,Wc", [(2, 12, 16), (1, 5, 7), (3, 9, 10), (1, 80, 80), (2, 23, 40)])
@pytest.mark.parametrize("variant", [0, 1])
def test_heatmap_nhwc_vector_path_equals_nchw(B, Hc, Wc, variant):
    """The engine's `semi` buffer is NHWC with 80-channel rows (float4 loads, paired-lane stores); the API path is NCHW (scalar
    loads).  Same arithmetic, so the heatmaps must be bit-identical -- including odd Wc (unpaired stores) and a partial last CTA."""
    g = torch.Generator().manual_seed(B * 100 + Wc)
    semi = (torch.randn(B, 65, Hc, Wc, generator=g) * 3).cuda()
    ref = ops.heatmap(semi, "nchw", variant=variant)
    padded = torch.full((B, Hc, Wc, 80), 1e30, device="cuda")          # channels 65..79 must never be read into the softmax
    padded[..., :65] = semi.permute(0, 2, 3, 1)
    got = ops.heatmap(padded, "nhwc", variant=variant)
    assert torch.equal(got, ref)
    if variant == 0:
        np.testing.assert_allclose(got.cpu().numpy(), O.flatten_detection(semi.cpu().numpy(), variant="torch"), rtol=0, atol=1e-6)


@pytest.mark.parametrize("B,ny,nx,nc,ldc", [(2, 5, 7, 80, 256), (1, 20, 20, 80, 256), (3, 4, 6, 1, 32), (1, 9, 3, 20, 80)])
@pytest.mark.parametrize("want_raw", [True, False])
def test_detect_decode_vs_oracle(B, ny, nx, nc, ldc, want_raw):
    L = _lib.lib(require_device=True)
    no, na = nc + 5, 3
    g = torch.Generator().manual_seed(ny * nx + nc)
    logits = torch.randn(B, ny, nx, ldc, generator=g) * 3
    anchors_px = torch.tensor([[10., 13.], [16., 30.], [33., 23.]])
    stride = 16.0
    A_tot, row_off = 3 * ny * nx + 11, 5                       # rows of another level around this one stay untouched
    pred = torch.full((B, A_tot, no), -7.0, device="cuda")
    raw = torch.full((B, na, ny, nx, no), -7.0, device="cuda") if want_raw else None
    dl = logits.cuda()
    anc = (C.c_float * 6)(*anchors_px.flatten().tolist())
    _lib.check(L.yp_detect_decode(dl.data_ptr(), B, ny, nx, ldc, na, no, stride, anc, raw.data_ptr() if want_raw else None, pred.data_ptr(), A_tot,
                                  row_off, _st()))
    torch.cuda.synchronize()
    raw_ref = logits[..., :na * no].view(B, ny, nx, na, no).permute(0, 3, 1, 2, 4).contiguous()
    ref = O.detect_decode([raw_ref], (anchors_px / stride)[None], torch.tensor([stride]))
    if want_raw:
        assert torch.equal(raw.cpu(), raw_ref)
    got = pred.cpu()
    assert float(got[:, :row_off].min()) == -7.0 and float(got[:, row_off + 3 * ny * nx:].max()) == -7.0
    np.testing.assert_allclose(got[:, row_off:row_off + 3 * ny * nx].numpy(), ref.numpy(), rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("fmt", [YP_FMT_BF16, YP_FMT_F32X2])
@pytest.mark.parametrize("B,H,W", [(2, 8, 36), (1, 6, 34), (1, 64, 96)])
def test_frame_to_s2d(fmt, B, H, W):
    """uint8 HWC frame -> 2x2 space-to-depth operand of the stem: W % 4 == 0 takes the two-pixel kernel, W = 34 the general one."""
    L = _lib.lib(require_device=True)
    frame = torch.from_numpy(np.random.RandomState(W).randint(0, 256, (B, H, W, 3)).astype(np.uint8)).cuda()
    planes, dt = (1, torch.bfloat16) if fmt == YP_FMT_BF16 else (2, torch.float32)
    out = torch.full((planes, B, H // 2, W // 2, 16), 9.0, dtype=dt, device="cuda")
    v = make_view(out, fmt)
    _lib.check(L.yp_frame_to_s2d(frame.data_ptr(), B, H, W, C.byref(v), _st()))
    torch.cuda.synchronize()
    x = torch.div(frame.float(), torch.full((), 255.0, device="cuda").expand(frame.shape))          # IEEE division (a python scalar divisor is a reciprocal multiply)
    want = torch.zeros(B, H // 2, W // 2, 16, device="cuda")
    want[..., :12] = x.view(B, H // 2, 2, W // 2, 2, 3).permute(0, 1, 3, 2, 4, 5).reshape(B, H // 2, W // 2, 12)
    if fmt == YP_FMT_BF16:
        assert torch.equal(out[0], want.to(torch.bfloat16))
    else:
        assert torch.equal(out, split_tf32(want))


@pytest.mark.parametrize("D", [64, 96, 128, 192, 256, 320])
def test_sample_desc_all_widths_and_out_of_range_points(D):
    """Every descriptor width of the model family (compile-time channel-block counts 2/4/6/8, run-time path for 96 / 320) against
    the oracle, with points left / right / above / below the image: out-of-bounds bilinear corners contribute zero (zeros padding
    of grid_sample, src/evaluations/descriptor_evaluation.py:166-176) and must not be read as data."""
    Hc, Wc, H, W = 12, 20, 96, 160
    rs = np.random.RandomState(D)
    coarse = rs.normal(0, 1, (1, D, Hc, Wc)).astype(np.float32)
    coarse /= np.linalg.norm(coarse, axis=1, keepdims=True)
    inside = np.stack((rs.randint(0, W, 300), rs.randint(0, H, 300)))
    edge = np.array([[-5, W + 5, 30, 40, -3, W + 4, W, 0], [20, 30, -4, H + 6, -2, H + 3, H, 0]])
    pts = np.concatenate((inside, edge), 1).astype(np.float64)
    pts = np.vstack((pts, rs.uniform(0, 1, pts.shape[1])))
    ref = O.sample_desc_from_points(coarse, pts, HW=(H, W))
    nhwc = torch.from_numpy(coarse).cuda().permute(0, 2, 3, 1).contiguous()
    p = torch.from_numpy(pts.T.astype(np.float32)).cuda()[None].contiguous()
    got = ops.sample_desc(nhwc, p, None, (H, W), "nhwc")[0].T.cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-7)
    got2 = ops.sample_desc(torch.from_numpy(coarse).cuda(), p, None, (H, W), "nchw")[0].T.cpu().numpy()      # strided (API) layout
    np.testing.assert_array_equal(got2, got)


@pytest.mark.parametrize("B,H,W,Ct,c_off,Cv,C", [(2, 5, 7, 80, 0, 80, 65), (1, 12, 20, 192, 0, 192, 192), (3, 9, 9, 96, 32, 64, 64), (1, 4, 8, 16, 0, 16, 16)])
def test_nhwc_to_nchw_export(B, H, W, Ct, c_off, Cv, C):
    """fp32 NHWC view (channel slice of a wider buffer, padded channel count) -> dense NCHW of the first C channels."""
    from yolopoint_b200._lib import YP_FMT_F32
    L = _lib.lib(require_device=True)
    t = torch.randn(1, B, H, W, Ct, device="cuda")
    v = make_view(t, YP_FMT_F32, c_off, Cv)
    out = torch.full((B, C, H, W), -3.0, device="cuda")
    _lib.check(L.yp_nhwc_to_nchw(ctypes.byref(v), C, out.data_ptr(), _st()))
    torch.cuda.synchronize()
    assert torch.equal(out, t[0, ..., c_off:c_off + C].permute(0, 3, 1, 2))


def test_l2norm_rows_and_tf32_split_kernels():
    """yp_l2norm_nhwc (descriptor normalisation for widths that do not fit one accumulator tile: version "x") against torch on a
    channel slice of a wider buffer, and yp_split_tf32 (operand planes of the tensor-core match) bit-exact against the PyTorch
    restatement of cvt.rna.tf32 (engine.split_tf32)."""
    import ctypes as C
    from yolopoint_b200 import _lib
    from yolopoint_b200._lib import YP_FMT_F32
    from yolopoint_b200.engine import make_view, split_tf32
    L = _lib.lib(require_device=True)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cpu").manual_seed(5)
    buf = torch.randn(1, 2, 9, 11, 320 + 16, generator=g).cuda()
    ref = buf.clone()
    ref[..., 16:] = ref[..., 16:] / ref[..., 16:].norm(dim=-1, keepdim=True)
    v = make_view(buf, YP_FMT_F32, 16, 320)
    _lib.check(L.yp_l2norm_nhwc(C.byref(v), st))
    torch.cuda.synchronize()
    assert torch.equal(buf[..., :16], ref[..., :16])                       # channels outside the slice untouched
    assert float((buf[..., 16:] - ref[..., 16:]).abs().max()) < 2e-7
    x = (torch.randn(1000, 256, generator=g) * torch.logspace(-6, 6, 1000).unsqueeze(1)).cuda()
    out = torch.empty((2, 1000, 256), device="cuda")
    _lib.check(L.yp_split_tf32(x.data_ptr(), x.numel(), out[0].data_ptr(), out[1].data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(out, split_tf32(x))


@pytest.mark.parametrize("fmt", ["f32x2", "bf16", "f32"])
@pytest.mark.parametrize("B,H,W,Cc", [(1, 20, 20, 256), (2, 15, 20, 32), (3, 4, 3, 8), (1, 23, 40, 16)])
def test_sppf_pool_equals_chained_maxpool(fmt, B, H, W, Cc):
    """yp_sppf_pool alone (the network tests only see it through whole passes): slices 1..3 of the [B,H,W,4C] concat buffer ==
    three chained nn.MaxPool2d(5, 1, 2) of slice 0 (src/models/common.py:220-229) on the stored values, for every activation
    format; the operand planes of the winning element are copied bit-exactly (distinct values, so the winner is unique), slice 0
    and the channels outside the view stay untouched.  Maps narrower than the window (4x3) exercise the clipped windows."""
    import torch.nn.functional as F
    from yolopoint_b200._lib import YP_FMT_F32
    L = _lib.lib(require_device=True)
    code = {"f32x2": YP_FMT_F32X2, "bf16": YP_FMT_BF16, "f32": YP_FMT_F32}[fmt]
    g = torch.Generator().manual_seed(17 * B + W + Cc)
    n = B * H * W * Cc
    x = (torch.randperm(n, generator=g).float() / n * 8 - 4).reshape(B, H, W, Cc).cuda()      # distinct values: unique winners
    planes, dt = (2, torch.float32) if fmt == "f32x2" else ((1, torch.bfloat16) if fmt == "bf16" else (1, torch.float32))
    buf = torch.full((planes, B, H, W, 4 * Cc + 16), 7.0, dtype=dt, device="cuda")
    if fmt == "f32x2":
        buf[:, ..., 16:16 + Cc] = split_tf32(x)
    else:
        buf[0, ..., 16:16 + Cc] = x.to(dt)
    before = buf.clone()
    v = make_view(buf, code, 16, 4 * Cc)
    _lib.check(L.yp_sppf_pool(C.byref(v), _st()))
    torch.cuda.synchronize()
    assert torch.equal(buf[..., :16 + Cc], before[..., :16 + Cc])
    val = before[..., 16:16 + Cc].float().sum(0).permute(0, 3, 1, 2)                            # the value the kernel sees (hi + lo)
    y = val
    for k in (1, 2, 3):
        y = F.max_pool2d(y, 5, 1, 2)
        got = buf[..., 16 + k * Cc:16 + (k + 1) * Cc]
        assert torch.equal(got.float().sum(0).permute(0, 3, 1, 2), y)
        if fmt != "bf16":                                                                       # bf16 rounding makes equal values; their bits are equal too
            w = 4 * k + 1
            _, idx = F.max_pool2d(val, w, 1, w // 2, return_indices=True)                       # chained 5x5 pools == one clipped (4k+1)^2 window
            for p in range(planes):
                flat = before[p, ..., 16:16 + Cc].permute(0, 3, 1, 2).reshape(B, Cc, H * W)
                want = torch.gather(flat, 2, idx.reshape(B, Cc, -1)).reshape(B, Cc, H, W)
                assert torch.equal(got[p].permute(0, 3, 1, 2), want)


def test_sppf_pool_rejects_bad_views():
    from yolopoint_b200._lib import YP_FMT_F32
    L = _lib.lib(require_device=True)
    buf = torch.zeros((1, 1, 8, 8, 24), device="cuda")
    assert L.yp_sppf_pool(C.byref(make_view(buf, YP_FMT_F32, 0, 6)), _st()) != 0       # 6 channels: not four slices of channel pairs
    assert L.yp_sppf_pool(None, _st()) != 0
    big = torch.zeros((1, 1, 300, 300, 8), device="cuda")
    assert L.yp_sppf_pool(C.byref(make_view(big, YP_FMT_F32, 0, 8)), _st()) != 0       # 90 000 pixels: beyond the 16-bit source index
