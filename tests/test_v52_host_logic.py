"""CPU-only tests of the YOLOPointv52 row (reference src/models/YOLOPoint.py:248-342): state-dict
compatibility with the reference, the oracle restatement against vectors the unmodified reference produced, and the launch plan
+ weight packing through the CPU plan interpreter."""
import numpy as np
import pytest
import torch

from oracle import yolopoint_oracle as O
from yolopoint_b200 import Model
from yolopoint_b200.engine import ConvOp, NetPlan, Pool2Op, PoolOp
from yolopoint_b200.synth import perturb_state_dict, synthetic_frame

from plan_interp import run_plan

NAMES = [str(i) for i in range(80)]
V52 = "YOLOPointv52"


@pytest.mark.parametrize("ver", ["n", "s"])
def test_state_dict_matches_reference(golden, ver):
    """Same keys, order, shapes and seeded values as the reference Model(model_name='YOLOPointv52')."""
    g = golden(f"state_v52{ver}.npz")
    torch.manual_seed(0)
    m = Model(names=NAMES, version=ver, model_name=V52)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(g["shapes"])
    np.testing.assert_allclose([float(v.double().sum()) for v in sd.values()], g["sums"], rtol=0, atol=1e-9)
    np.testing.assert_allclose([float(v.double().abs().sum()) for v in sd.values()], g["asums"], rtol=0, atol=1e-9)
    assert [n for n, _ in m.named_parameters()] == list(g["param_names"])
    np.testing.assert_array_equal(m.model.Detect.stride.numpy(), g["stride"])
    np.testing.assert_array_equal(m.model.Detect.anchors.numpy(), g["anchors"])


def _golden_net(golden):
    g = golden("net_v52n_64x96.npz")
    torch.manual_seed(0)
    m = Model(names=NAMES, version="n", model_name=V52)
    sd = perturb_state_dict(m.state_dict(), 0, "n")
    return g, m, sd


def test_oracle_network_matches_reference_golden(golden):
    """OracleNet.forward_v52 against the reference's fused eval forward (same torch-CPU kernels: equal to fp32 rounding)."""
    g, _, sd = _golden_net(golden)
    o = O.OracleNet(sd, "n", 80, V52).forward(torch.from_numpy(g["x"]))
    np.testing.assert_allclose(o["semi"].numpy(), g["semi"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(o["desc"].numpy(), g["desc"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(o["objects"][0].numpy(), g["pred"], rtol=1e-6, atol=1e-5)
    for i in range(3):
        np.testing.assert_allclose(o["objects"][1][i].numpy(), g[f"raw{i}"], rtol=0, atol=1e-5)


def test_torch_training_tree_matches_golden_in_eval_math(golden):
    g, m, sd = _golden_net(golden)
    m.load_state_dict(sd)
    m.model.eval()
    with torch.no_grad():
        o = m.model(torch.from_numpy(g["x"]))
    np.testing.assert_allclose(o["semi"].numpy(), g["semi"], rtol=0, atol=5e-4)
    np.testing.assert_allclose(o["desc"].numpy(), g["desc"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(o["objects"][0].numpy(), g["pred"], rtol=1e-5, atol=2e-3)


def test_oracle_whole_frame_matches_reference_golden(golden):
    """process_frame on YOLOPointv52-S 640x640 equals what YoloPointFrontend.process_img of the reference returned."""
    g = golden("e2e_v52s_640x640.npz")
    torch.manual_seed(0)
    m = Model(names=NAMES, version="s", model_name=V52)
    net = O.OracleNet(perturb_state_dict(m.state_dict(), 0, "s"), "s", 80, V52)
    for i in (0, 1):
        pts, desc, boxes = O.process_frame(net, synthetic_frame(640, 640, i))
        np.testing.assert_array_equal(pts, g[f"pts{i}"])
        np.testing.assert_array_equal(boxes, g[f"boxes{i}"])
        np.testing.assert_allclose(desc, g[f"desc{i}"], rtol=0, atol=1e-6)
    # the match on the reference's own descriptors (1e-7 descriptor noise may flip a pair that sits on the nn_thresh boundary)
    np.testing.assert_array_equal(O.nn_match_two_way(g["desc0"], g["desc1"], O.DEFAULT_CFG["nn_thresh"]), g["matches"])


@pytest.mark.parametrize("ver,shape", [("n", (2, 64, 96)), ("s", (1, 64, 64)), ("m", (1, 64, 64))])
def test_plan_reproduces_oracle(ver, shape):
    """Launch plan (C2f chunk / cat as channel slices, 2x2 max-pool into the concat buffer, two-destination C2f / SPPF outputs, BN +
    SiLU heads with the L2-norm epilogue) + packed weights, interpreted on the CPU, equal the oracle network."""
    torch.manual_seed(0)
    m = Model(names=NAMES, version=ver, model_name=V52)
    sd = perturb_state_dict(m.state_dict(), 0, ver)
    B, H, W = shape
    x = torch.from_numpy(np.random.RandomState(3).rand(B, 3, H, W).astype(np.float32))
    ref = O.OracleNet(sd, ver, 80, V52).forward(x)
    net = NetPlan(ver, 80, "fp32", V52)
    bufs = run_plan(net, sd, x)
    np.testing.assert_allclose(bufs["semi"][..., :65].permute(0, 3, 1, 2).numpy(), ref["semi"].numpy(), rtol=0, atol=2e-4)
    assert float(bufs["semi"][..., 65:].abs().max()) == 0.0          # padded channels: zero weights, zero bias, SiLU(0) = 0
    np.testing.assert_allclose(bufs["desc"].permute(0, 3, 1, 2).numpy(), ref["desc"].numpy(), rtol=0, atol=2e-5)
    for i in range(3):
        det = bufs[f"det{i}"][..., :255]
        raw = det.view(B, det.shape[1], det.shape[2], 3, 85).permute(0, 3, 1, 2, 4)
        np.testing.assert_allclose(raw.numpy(), ref["objects"][1][i].numpy(), rtol=0, atol=5e-4)


@pytest.mark.parametrize("ver", ["n", "s", "m", "l"])
def test_plan_structure(ver):
    """Every view the plan hands to the kernels satisfies the C-ABI contract (16-channel granularity of slices, one writer per
    channel slice and pass) and the launch count / FLOPs are the reference's."""
    net = NetPlan(ver, 80, "bf16", V52)
    n_conv_ref = {"n": 59, "s": 59, "m": 85, "l": 111}[ver]          # conv modules of the reference model (hook count)
    assert len(net.conv_ops()) == n_conv_ref                           # C2f has no cv1 || cv2 pair to merge
    assert sum(isinstance(op, Pool2Op) for op in net.ops) == 1 and sum(isinstance(op, PoolOp) for op in net.ops) == 1
    for op in net.ops:
        refs = [op.src, *op.dst] if isinstance(op, ConvOp) else ([op.src, op.dst] if isinstance(op, Pool2Op) else [])
        for r in refs:
            _, ctot, _ = net.bufs[r.buf]
            assert r.c_off % 16 == 0 and r.C % 16 == 0 and r.c_off + r.C <= ctot, (op, r)
        if isinstance(op, ConvOp):
            # a launch never reads the slice it writes (no residual in C2f; the bottlenecks go through the .h buffer)
            for d in op.dst:
                assert not (d.buf == op.src.buf and d.c_off < op.src.c_off + op.src.C and op.src.c_off < d.c_off + d.C), op
            assert op.residual is None
    if ver == "s":
        torch.manual_seed(0)
        sd = Model(names=NAMES, version="s", model_name=V52).state_dict()
        meta = {}
        for op in net.conv_ops():
            for n in op.names:
                w = sd[f"model.{n}.conv.weight" if op.bn else f"model.{n}.weight"]
                meta[n] = (w.shape[0], w.shape[1], w.shape[2])
        assert abs(net.flops_per_frame(640, 640, meta) / 1e9 - 21.363) < 0.01


def test_unknown_model_name_raises():
    with pytest.raises(NotImplementedError):
        Model(names=NAMES, version="n", model_name="YOLOPointM")


def test_fuse_partial_load_and_cpu_training_path():
    """Wrapper behaviour for the v52 tree: fuse() folds every Conv block's BN (src/models/YOLOPoint.py:84-90), a checkpoint with another
    class count loads positionally except Detect (:102-135), and the train-mode module tree backpropagates to every parameter on the
    CPU (plain PyTorch path: the reference's own way of running it)."""
    torch.manual_seed(0)
    m = Model(names=NAMES, version="n", model_name=V52)
    sd = m.state_dict()
    m2 = Model(names=["a", "b", "c"], version="n", model_name=V52)
    m2.load_state_dict(sd, strict=True)
    assert torch.equal(m2.state_dict()["model.Bottleneck1.m.0.cv2.conv.weight"], sd["model.Bottleneck1.m.0.cv2.conv.weight"])
    assert m2.state_dict()["model.Detect.m.0.bias"].shape[0] == 3 * 8
    m.train()
    x = torch.rand(2, 3, 64, 96)
    out = m(x)
    assert out["semi"].shape == (2, 65, 8, 12) and out["desc"].shape == (2, 64, 8, 12) and len(out["objects"]) == 3
    loss = out["semi"].square().mean() + out["desc"][:, ::2].mean() + sum(r.square().mean() for r in out["objects"])
    loss.backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters())
    assert len(list(m.parameters())) == 174
    m.eval().fuse()
    keys = list(m.state_dict().keys())
    assert "model.BottleneckDet.cv2.conv.bias" in keys and not any(".bn." in k for k in keys)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))      # eval-mode inference on CPU must fail loudly


def test_synthetic_weights_and_tuning_tables_are_model_specific():
    """perturb_state_dict detects the v52 key set (no ConvDet) and peaks the keypoint logits through BottleneckDet.cv2's BN weight;
    tuning tables of the two model families never alias (a YOLOPoint table names layers of other shapes)."""
    from yolopoint_b200.engine import load_tuning, tuning_path
    from yolopoint_b200.synth import TUNING, TUNING_V52, V52_SEMI_GAIN
    torch.manual_seed(0)
    sd = Model(names=NAMES, version="n", model_name=V52).state_dict()
    a, b = perturb_state_dict(sd, 0, "n"), perturb_state_dict(sd, 0, "n")
    assert all(torch.equal(a[k], b[k]) for k in a)                                   # deterministic
    plain = perturb_state_dict(sd, 0, "n", semi_gain=1.0)
    k = "model.BottleneckDet.cv2.bn.weight"
    assert torch.allclose(a[k], plain[k] * V52_SEMI_GAIN) and all(torch.equal(a[j], plain[j]) for j in a if j != k)
    assert TUNING_V52["n"] != TUNING["n"] and TUNING_V52["m"] == TUNING["m"]
    p5, p52 = tuning_path("s", 1, 640, 640, "fp32"), tuning_path("s", 1, 640, 640, "fp32", V52)
    assert p5 != p52 and p52.endswith("v52s_1x640x640_fp32.json")
    t52, t5 = load_tuning("s", 1, 640, 640, "fp32", V52), load_tuning("s", 1, 640, 640, "fp32")
    assert t52 and "BottleneckDet.cv2" in t52 and "ConvDet" not in t52 and "ConvDet" in t5   # each family has its own measured table
    assert load_tuning("l", 1, 640, 640, "fp32", V52) == {}                          # no table: library heuristics
    assert len(load_tuning("s", 1, 640, 640, "fp32")) > 0
