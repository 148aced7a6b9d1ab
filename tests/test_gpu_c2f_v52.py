"""GPU parity of the YOLOPointv52 row (reference src/models/YOLOPoint.py:248-342): the 2x2 max-pool
kernel, the conv geometries only this model has, the network through the engine against the oracle / the reference golden, the
whole-frame pipeline against what the unmodified reference produced, and one training step on the tcgen05 kernels.
Tolerances as in test_gpu_network.py."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import yolopoint_b200 as yp
from oracle import yolopoint_oracle as O
from yolopoint_b200 import FramePipeline, Model, _lib
from yolopoint_b200._lib import YP_ALGO_TCGEN05, YP_FMT_BF16, YP_FMT_F32, YP_FMT_F32X2
from yolopoint_b200.engine import make_view, split_tf32
from yolopoint_b200.synth import perturb_state_dict, synthetic_frame

from test_gpu_conv import run_case
from test_gpu_network import check_outputs

pytestmark = pytest.mark.gpu
NAMES = [str(i) for i in range(80)]
V52 = "YOLOPointv52"
_cache = {}


def build(ver, precision="fp32"):
    key = (ver, precision)
    if key not in _cache:
        torch.manual_seed(0)
        m = Model(names=NAMES, version=ver, precision=precision, model_name=V52)
        sd = perturb_state_dict(m.state_dict(), 0, ver)
        m.load_state_dict(sd)
        _cache[key] = (m.cuda().eval(), sd)
    return _cache[key]


@pytest.mark.parametrize("fmt", [YP_FMT_F32X2, YP_FMT_BF16, YP_FMT_F32])
@pytest.mark.parametrize("B,H,W,Cc", [(1, 160, 160, 64), (2, 30, 44, 16), (3, 8, 8, 96)])
def test_maxpool2x2(fmt, B, H, W, Cc):
    """yp_maxpool2x2 from a channel slice into a channel slice of a concat buffer == F.max_pool2d(2, 2) on the stored values,
    the winner's operand planes copied bit-exactly, neighbouring channels untouched."""
    L = _lib.lib(require_device=True)
    g = torch.Generator().manual_seed(B * H + Cc)
    x = torch.randn(B, H, W, Cc, generator=g).cuda()
    planes, dt = (2, torch.float32) if fmt == YP_FMT_F32X2 else ((1, torch.bfloat16) if fmt == YP_FMT_BF16 else (1, torch.float32))
    src = torch.full((planes, B, H, W, Cc + 16), 3.0, dtype=dt, device="cuda")
    if fmt == YP_FMT_F32X2:
        src[:, ..., 16:] = split_tf32(x)
    else:
        src[0, ..., 16:] = x.to(dt)
    dst = torch.full((planes, B, H // 2, W // 2, 2 * Cc + 32), 5.0, dtype=dt, device="cuda")
    vi, vo = make_view(src, fmt, 16, Cc), make_view(dst, fmt, 32, Cc)
    _lib.check(L.yp_maxpool2x2(C.byref(vi), C.byref(vo), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    val = src[..., 16:].float().sum(0)                                    # the value the kernels see (hi + lo)
    ref, idx = F.max_pool2d(val.permute(0, 3, 1, 2), 2, 2, return_indices=True)
    got = dst[..., 32:32 + Cc].float().sum(0).permute(0, 3, 1, 2)
    assert torch.equal(got, ref)
    for p in range(planes):                                               # planes of the winning element, unchanged
        flat = src[p, ..., 16:].permute(0, 3, 1, 2).reshape(B, Cc, H * W)
        want = torch.gather(flat, 2, idx.reshape(B, Cc, -1)).reshape(B, Cc, H // 2, W // 2)
        assert torch.equal(dst[p, ..., 32:32 + Cc].permute(0, 3, 1, 2), want)
    rest = torch.cat((dst[..., :32], dst[..., 32 + Cc:]), -1).float()
    assert float(rest.min()) == 5.0 and float(rest.max()) == 5.0


def test_maxpool2x2_rejects_bad_geometry():
    L = _lib.lib(require_device=True)
    a = torch.zeros((1, 1, 6, 6, 16), dtype=torch.bfloat16, device="cuda")
    b = torch.zeros((1, 1, 2, 3, 16), dtype=torch.bfloat16, device="cuda")
    rc = L.yp_maxpool2x2(C.byref(make_view(a, YP_FMT_BF16)), C.byref(make_view(b, YP_FMT_BF16)), None)
    assert rc != 0 and b"maxpool2x2" in L.yp_last_error()


V52_CONV_CASES = [
    dict(B=1, H=80, W=80, Cin=192, Cout=128, k=1, s=1, l2=True, plain=True),     # BottleneckDesc.cv2 (S): BN bias + SiLU, then L2 norm
    dict(B=1, H=80, W=80, Cin=96, Cout=80, k=1, s=1, plain=True),                # BottleneckDet.cv2: 65 (padded 80) SiLU outputs in fp32
    dict(B=1, H=120, W=160, Cin=48, Cout=32, k=1, s=1),                          # Bottleneck1.cv2 (N): (2+n)c = 48 input channels
    dict(B=1, H=160, W=160, Cin=32, Cout=32, k=3, s=1),                          # Bottleneckv8 3x3 on a chunk of the C2f buffer (S)
    dict(B=1, H=40, W=40, Cin=384, Cout=256, k=1, s=1, up=True),                 # Bottleneck5.cv2 (S): xe to cat7 and 2x upsampled to cat6
    dict(B=2, H=46, W=80, Cin=144, Cout=96, k=1, s=1),                           # Bottleneck1.cv2 (M): 3 x 48 channels
    dict(B=1, H=80, W=80, Cin=288, Cout=192, k=1, s=1, l2=True, plain=True),     # BottleneckDesc.cv2 (M): 192-wide L2 norm
]


@pytest.mark.parametrize("ci", range(len(V52_CONV_CASES)))
@pytest.mark.parametrize("fmt", [YP_FMT_F32X2, YP_FMT_BF16])
def test_conv_geometries_of_v52(ci, fmt):
    run_case(V52_CONV_CASES[ci], fmt, YP_ALGO_TCGEN05)


def test_forward_golden_n(golden):
    g = golden("net_v52n_64x96.npz")
    m, _ = build("n")
    out = m(torch.from_numpy(g["x"]).cuda())
    ref = dict(semi=torch.from_numpy(g["semi"]), desc=torch.from_numpy(g["desc"]),
               objects=(torch.from_numpy(g["pred"]), [torch.from_numpy(g[f"raw{i}"]) for i in range(3)]))
    check_outputs(out, ref, "v52 golden n 64x96")


@pytest.mark.parametrize("ver,B,H,W", [("n", 1, 480, 640), ("s", 1, 640, 640), ("s", 3, 96, 160), ("m", 1, 128, 160), ("l", 1, 64, 96)])
def test_forward_vs_oracle(ver, B, H, W):
    m, sd = build(ver)
    x = torch.from_numpy(np.random.RandomState(H + W).rand(B, 3, H, W).astype(np.float32))
    out = m(x.cuda())
    ref = O.OracleNet(sd, ver, 80, V52).forward(x)
    check_outputs(out, ref, f"v52 {ver} {B}x{H}x{W}")
    out2 = m(x.cuda())   # second call replays the CUDA graph: must be identical
    assert torch.equal(out["semi"], out2["semi"]) and torch.equal(out["objects"][0], out2["objects"][0]) and torch.equal(out["desc"], out2["desc"])


def test_forward_bf16_mode_is_close():
    m, sd = build("s", "bf16")
    x = torch.from_numpy(np.random.RandomState(1).rand(2, 3, 256, 384).astype(np.float32))
    out = m(x.cuda())
    ref = O.OracleNet(sd, "s", 80, V52).forward(x)
    e = float((out["desc"].cpu() - ref["desc"]).abs().max())
    es = float((out["semi"].cpu() - ref["semi"]).abs().max()) / float(ref["semi"].abs().max())
    print("v52 bf16 desc max abs err", e, "semi rel", es)
    assert e < 5e-2 and es < 5e-2


def test_pipeline_stages_bit_exact_on_same_inputs():
    """The GPU network's own outputs into both the kernels and the oracle post-processing: indices bit-exact."""
    m, sd = build("s")
    H = W = 640
    frame = synthetic_frame(H, W, 0)
    x = torch.from_numpy(frame.transpose(2, 0, 1).astype(np.float32) / 255.)[None]
    out = m(x.cuda())
    cfg = O.DEFAULT_CFG
    outs_cpu = dict(semi=out["semi"].cpu(), desc=out["desc"].cpu(), objects=(out["objects"][0].cpu(), None))
    pts_ref, desc_ref, boxes_ref = O.process_outputs(outs_cpu, H, W, cfg, True, heat_variant="torch")
    boxes = yp.non_max_suppression(out["objects"][0], cfg["conf_thres_box"], cfg["iou_thres_box"], multi_label=True, agnostic=True,
                                   max_det=cfg["max_det"])[0]
    np.testing.assert_array_equal(boxes.cpu().numpy(), boxes_ref)
    pts, desc = yp.extract_keypoints(out["semi"], out["desc"], cfg["detection_threshold"], cfg["nms"], boxes=boxes)
    heat_ref = O.flatten_detection(outs_cpu["semi"].numpy()[0])
    near = np.abs(heat_ref - cfg["detection_threshold"]) < 1e-6
    print("v52: pixels within 1e-6 of the detection threshold:", int(near.sum()), "keypoints", pts.shape[1], "boxes", boxes.shape[0])
    if near.sum() == 0:
        assert pts.shape == pts_ref.shape
        np.testing.assert_array_equal(pts[:2], pts_ref[:2])
        np.testing.assert_allclose(pts[2], pts_ref[2], rtol=0, atol=1e-6)
        np.testing.assert_allclose(desc, desc_ref, rtol=0, atol=1e-5)


def test_frame_pipeline_vs_golden(golden):
    """Whole-frame pipeline from uint8 host frames against what the unmodified reference (YOLOPointv52-S) produced."""
    H = W = 640
    g = golden("e2e_v52s_640x640.npz")
    m, _ = build("s")
    pipe = FramePipeline(m, 1, H, W)
    res = [pipe.step_host(synthetic_frame(H, W, s)[None])[0] for s in (0, 1)]
    for i, (pts, desc, boxes, matches) in enumerate(res):
        rp, rd, rb = g[f"pts{i}"], g[f"desc{i}"], g[f"boxes{i}"]
        ref_set = {(int(x), int(y)) for x, y in zip(rp[0], rp[1])}
        got_set = {(int(x), int(y)) for x, y in zip(pts[0], pts[1])}
        common = len(ref_set & got_set)
        print(f"v52 s frame {i}: keypoints ref {len(ref_set)} got {len(got_set)} common {common}; boxes ref {rb.shape[0]} got {boxes.shape[0]}")
        assert common >= 0.99 * len(ref_set) and len(got_set) <= 1.01 * len(ref_set) + 1
        if got_set == ref_set and pts.shape == rp.shape and np.array_equal(pts[:2], rp[:2]):
            np.testing.assert_allclose(pts[2], rp[2], rtol=0, atol=1e-5)
            np.testing.assert_allclose(desc, rd, rtol=0, atol=1e-4)
        assert abs(boxes.shape[0] - rb.shape[0]) <= max(2, rb.shape[0] // 20)
        if boxes.shape == rb.shape:
            d = np.abs(boxes[None, :, :4] - rb[:, None, :4]).max(-1).min(1)
            assert (d < 0.5).mean() >= 0.95, float((d < 0.5).mean())
    rm = g["matches"]
    print(f"v52 s: matches ref {rm.shape[1]} got {res[1][3].shape[1]}")
    assert abs(res[1][3].shape[1] - rm.shape[1]) <= max(3, rm.shape[1] // 20)


def test_model_train_step_on_tc_kernels():
    """One train-mode forward/backward of YOLOPointv52-N with the convolutions (forward, data and weight gradients) on the tcgen05
    kernels vs cuDNN on the same bf16 channels-last dataflow: same loss, every parameter gets a finite gradient, the layers next
    to the loss agree closely."""
    torch.manual_seed(0)
    m = Model(names=NAMES, version="n", model_name=V52).cuda().train()
    x = torch.rand(4, 3, 192, 256, device="cuda")
    state = {k: v.clone() for k, v in m.state_dict().items()}
    res = {}
    for backend in ("cudnn_bf16", "b200"):
        m.load_state_dict(state)
        m.train_backend = backend
        m.zero_grad(set_to_none=True)
        out = m(x)
        loss = out["semi"].float().square().mean() + out["desc"].float()[:, ::2].mean() * 3 + sum(r.float().square().mean() for r in out["objects"])
        loss.backward()
        res[backend] = (float(loss), {n: p.grad.detach().float().clone() for n, p in m.named_parameters() if p.grad is not None})
    assert any(type(mod).__name__ == "TcConv2d" and not mod._tc_cudnn for mod in m.modules())
    (loss_c, g_c), (loss_b, g_b) = res["cudnn_bf16"], res["b200"]
    assert len(g_b) == len(g_c) == len(list(m.parameters()))
    assert all(bool(torch.isfinite(v).all()) for v in g_b.values())
    assert abs(loss_b - loss_c) < 2e-2 * max(1.0, abs(loss_c)), (loss_b, loss_c)
    cs = {}
    for n in g_c:
        a, b = g_b[n].flatten().double(), g_c[n].flatten().double()
        if float(b.norm()) >= 1e-12:
            cs[n] = float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))
    c = np.array(list(cs.values()))
    heads = [v for n, v in cs.items() if n.startswith("model.Detect")]
    print(f"v52 grad cosine b200 vs cudnn_bf16: min {c.min():.4f} median {np.median(c):.4f} n {len(c)}; Detect layers min {min(heads):.5f}")
    assert min(heads) > 0.99 and np.median(c) > 0.5
