"""CPU-only tests of the host side: state-dict compatibility with the reference, launch plan + weight packing
(through the CPU plan interpreter), C-ABI export table."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import yolopoint_oracle as O
from yolopoint_b200 import Model, _lib
from yolopoint_b200.engine import NetPlan, split_tf32, tf32_round
from yolopoint_b200.synth import perturb_state_dict

from plan_interp import run_plan

NAMES = [str(i) for i in range(80)]


@pytest.mark.parametrize("ver", ["n", "s"])
def test_state_dict_matches_reference(golden, ver):
    """Same keys, order, shapes and seeded values as the reference Model (src/models/YOLOPoint.py:17-100)."""
    g = golden(f"state_{ver}.npz")
    torch.manual_seed(0)
    m = Model(names=NAMES, version=ver)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(g["shapes"])
    np.testing.assert_allclose([float(v.double().sum()) for v in sd.values()], g["sums"], rtol=0, atol=1e-9)
    np.testing.assert_allclose([float(v.double().abs().sum()) for v in sd.values()], g["asums"], rtol=0, atol=1e-9)
    assert [n for n, _ in m.named_parameters()] == list(g["param_names"])
    np.testing.assert_array_equal(m.model.Detect.stride.numpy(), g["stride"])
    np.testing.assert_array_equal(m.model.Detect.anchors.numpy(), g["anchors"])


def test_torch_training_path_matches_oracle_in_eval_math(golden):
    """The PyTorch module tree (used for training) computes the reference network: compare its eval math to the golden."""
    g = golden("net_n_64x96.npz")
    torch.manual_seed(0)
    m = Model(names=NAMES, version="n")
    m.load_state_dict(perturb_state_dict(m.state_dict(), 0, "n"))
    m.model.eval()
    with torch.no_grad():
        o = m.model(torch.from_numpy(g["x"]))
    np.testing.assert_allclose(o["semi"].numpy(), g["semi"], rtol=0, atol=5e-4)
    np.testing.assert_allclose(o["desc"].numpy(), g["desc"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(o["objects"][0].numpy(), g["pred"], rtol=1e-5, atol=5e-4)


def test_fuse_and_partial_load():
    torch.manual_seed(0)
    m = Model(names=NAMES, version="n")
    sd = m.state_dict()
    m2 = Model(names=["a", "b", "c"], version="n")
    m2.load_state_dict(sd, strict=True)   # class count changed -> positional partial load, Detect kept
    assert torch.equal(m2.state_dict()["model.Conv1.conv.weight"], sd["model.Conv1.conv.weight"])
    assert m2.state_dict()["model.Detect.m.0.bias"].shape[0] == 3 * 8
    m.eval().fuse()
    assert "model.Conv1.conv.bias" in m.state_dict() and "model.Conv1.bn.weight" not in m.state_dict()
    m.freeze_layers([0, 1], verbose=False)
    assert not list(m.parameters())[0].requires_grad
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))      # eval-mode inference on CPU must fail loudly


def test_tf32_split_is_exact_enough():
    x = torch.randn(10000) * 3
    hi = tf32_round(x)
    assert (hi.view(torch.int32) & 0x1FFF).abs().max() == 0
    s = split_tf32(x)
    assert ((s[0].double() + s[1].double() - x.double()).abs() / x.double().abs().clamp_min(1e-30)).max() < 2 ** -21


@pytest.mark.parametrize("ver,shape", [("n", (2, 64, 96)), ("s", (1, 64, 64))])
def test_plan_reproduces_oracle(ver, shape):
    """Launch plan + packed weights, interpreted on the CPU, equal the oracle network."""
    torch.manual_seed(0)
    m = Model(names=NAMES, version=ver)
    sd = perturb_state_dict(m.state_dict(), 0, ver)
    B, H, W = shape
    x = torch.from_numpy(np.random.RandomState(3).rand(B, 3, H, W).astype(np.float32))
    ref = O.OracleNet(sd, ver, 80).forward(x)
    net = NetPlan(ver, 80, "fp32")
    bufs = run_plan(net, sd, x)
    semi = bufs["semi"][..., :65].permute(0, 3, 1, 2)
    desc = bufs["desc"].permute(0, 3, 1, 2)
    np.testing.assert_allclose(semi.numpy(), ref["semi"].numpy(), rtol=0, atol=2e-4)
    np.testing.assert_allclose(desc.numpy(), ref["desc"].numpy(), rtol=0, atol=2e-5)
    for i in range(3):
        det = bufs[f"det{i}"][..., :255]
        raw = det.view(B, det.shape[1], det.shape[2], 3, 85).permute(0, 3, 1, 2, 4)
        np.testing.assert_allclose(raw.numpy(), ref["objects"][1][i].numpy(), rtol=0, atol=5e-4)


def test_plan_launch_count_and_flops():
    net = NetPlan("s", 80, "fp32")
    assert len(net.conv_ops()) == 74 - 10          # 10 C3 blocks each merge cv1 || cv2 into one GEMM
    torch.manual_seed(0)
    m = Model(names=NAMES, version="s")
    meta = {}
    for op in net.conv_ops():
        for n in op.names:
            key = f"model.{n}.conv.weight" if op.bn else f"model.{n}.weight"
            w = m.state_dict()[key]
            meta[n] = (w.shape[0], w.shape[1], w.shape[2])
    gf = net.flops_per_frame(640, 640, meta) / 1e9
    assert abs(gf - 21.023) < 0.01, gf               # 21.023 GFLOP / frame for S @ 640x640


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads (no GPU needed) and exports exactly the symbols include/yolopoint_b200.h declares."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "yolopoint_b200.h")).read()
    declared = set(re.findall(r"\b(yp_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.yp_abi_version() == 9
    # struct layouts must agree with the header (sizes only; offsets follow from the C rules both sides use)
    assert ctypes.sizeof(_lib.YpView) == 48 and ctypes.sizeof(_lib.YpNmsParams) == 40


def test_api_argument_errors_and_loud_failure_without_gpu():
    """The reference-named functions refuse bad arguments with the reference's exception types BEFORE touching the device
    (src/demo.py:318-321: ValueError for a negative nn_thresh, assert on the descriptor widths; src/utils/general_yolo.py:146-147:
    asserts on the thresholds), return the reference's empty result for empty descriptor sets (src/demo.py:316-317), and raise a
    RuntimeError -- never compute on the CPU -- when asked for real work without a CUDA device."""
    import yolopoint_b200 as yp
    if torch.cuda.is_available():
        pytest.skip("checks the behaviour of a machine without a GPU")
    d = np.random.RandomState(0).normal(0, 1, (32, 10)).astype(np.float32)
    with pytest.raises(ValueError):
        yp.nn_match_two_way(d, d, -1.0)
    with pytest.raises(AssertionError):
        yp.nn_match_two_way(d, d[:16], 0.7)
    with pytest.raises(AssertionError):
        yp.non_max_suppression(torch.zeros(1, 10, 6), 1.5, 0.45)
    with pytest.raises(AssertionError):
        yp.non_max_suppression(torch.zeros(1, 10, 6), 0.4, -0.1)
    m = yp.nn_match_two_way(np.zeros((32, 0), np.float32), d, 0.7)
    assert m.shape == (3, 0) and m.dtype == np.float64
    for call in (lambda: yp.nn_match_two_way(d, d, 0.7), lambda: yp.non_max_suppression(torch.rand(1, 10, 6), 0.4, 0.45),
                 lambda: yp.flattenDetection(torch.rand(1, 65, 4, 4)), lambda: yp.getPtsFromHeatmap(np.zeros((32, 32), np.float32), 0.1, 4),
                 lambda: yp.sample_desc_from_points(torch.rand(1, 32, 4, 4), np.zeros((3, 0)))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps every exported entry point to the reference code it stands in for (or says it has no counterpart)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    missing = [name for name in _lib.SIGNATURES if name not in doc]
    assert not missing, missing


def test_frontend_host_side_without_gpu():
    """preprocess / restore_coords / template_filter of YoloPointFrontend are pure host code (src/demo.py:97-123, 187-195, 217-228)."""
    import yolopoint_b200 as yp
    rs = np.random.RandomState(0)
    fe = yp.YoloPointFrontend(None)
    frame = rs.randint(0, 256, (490, 651, 3)).astype(np.uint8)
    img, cth, ctw, fac = fe.preprocess(frame)
    assert img.shape == (480, 640, 3) and (cth, ctw, fac) == (5, 6, 1.0)          # ceil(10/2), ceil(11/2)
    np.testing.assert_array_equal(img, frame[5:485, 6:646])
    fe2 = yp.YoloPointFrontend(None, {"crop_resize": [10, 410, 20, 820, 640], "model": {"superpoint": {"nms": 4}}})
    assert fe2.cfg["nms"] == 4
    img2, cth2, ctw2, fac2 = fe2.preprocess(rs.randint(0, 256, (480, 960, 3)).astype(np.uint8))
    assert img2.shape == (320, 640, 3) and fac2 == 0.8 and (cth2, ctw2) == (0, 0)   # 480x960 is a multiple of 32: no second crop
    pts = np.array([[10., 20., 30.], [40., 50., 60.], [.9, .8, .7]])
    boxes = torch.tensor([[8., 16., 24., 32., .5, 1.]])
    p, b = fe2.restore_coords(pts.copy(), boxes.clone(), 0, 0, 0.8)
    # x / y rows divided by the resize factor; then the reference's crop offsets land on the first two POINTS (all rows)
    np.testing.assert_allclose(p, np.array([[12.5 + 20, 25. + 10, 37.5], [50. + 20, 62.5 + 10, 75.], [.9 + 20, .8 + 10, .7]]))
    np.testing.assert_allclose(b[0, :4].numpy(), np.array([10. + 20, 20. + 10, 30. + 20, 40. + 10]))
    fe.templates["front"] = np.ones((480, 640))
    fe.templates["front"][:, 320:] = 0
    q = np.array([[100., 400., 319., 320.], [10., 20., 30., 40.], [.9, .8, .7, .6]])
    d = np.arange(8, dtype=np.float32).reshape(2, 4)
    fp, fd, dropped = fe.template_filter(q, d, "front")
    np.testing.assert_array_equal(fp, q[:, [0, 2]])
    np.testing.assert_array_equal(fd, d[:, [0, 2]])
    assert dropped


def test_load_model_signature_and_compat_on_stand_in_modules():
    """load_model keeps the reference's calling convention (src/utils/utils.py:55-57); compat.install() rebinds whatever of the
    reference's modules is importable -- exercised here on stand-in modules, against the live reference in test_oracle_vs_reference.py."""
    import sys
    import types
    import yolopoint_b200 as yp
    import yolopoint_b200.compat as compat
    m = yp.load_model(inp_ch=3, names=NAMES, version="n", model_name="YOLOPointv52")
    assert isinstance(m, Model) and m.model_name == "YOLOPointv52"
    with pytest.raises(NotImplementedError):
        yp.load_model(meta_model=False, model_name="SuperPointNet")
    fake_models, fake_utils = types.ModuleType("models"), types.ModuleType("utils.utils")
    fake_models.Model = object
    fake_utils.nms_fast = len
    saved = {k: sys.modules.get(k) for k in ("models", "utils.utils")}
    sys.modules["models"], sys.modules["utils.utils"] = fake_models, fake_utils
    try:
        done = compat.install(modules=("models", "utils.utils"))
        assert done == ["models.Model", "utils.utils.nms_fast"] and fake_models.Model is yp.Model and fake_utils.nms_fast is yp.nms_fast
        compat.uninstall()
        assert fake_models.Model is object and fake_utils.nms_fast is len
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_chain_dependencies_and_schedule():
    """Layer-chain host logic (engine.chain_dependencies / chain_schedule): every read-after-write, write-after-read and
    write-after-write pair of the launch list is ordered by the (transitively reduced) dependency lists, the schedule is a valid
    sequential order, and the segment cuts sit right behind the milestone layers."""
    from yolopoint_b200.engine import ConvOp, NetPlan, _op_reads_writes, _overlap, chain_dependencies, chain_schedule
    for model_name, ver in (("YOLOPoint", "s"), ("YOLOPoint", "l"), ("YOLOPointv52", "n")):
        net = NetPlan(ver, 80, "fp32", model_name)
        ops = net.ops
        deps = chain_dependencies(ops)
        # transitive closure of the reduced lists
        anc = []
        for j, d in enumerate(deps):
            a = set(d)
            for i in d:
                assert i < j
                a |= anc[i]
            anc.append(a)
        rw = [_op_reads_writes(op) for op in ops]
        for j in range(len(ops)):
            for i in range(j):
                hazard = _overlap(rw[i][1], rw[j][0]) or _overlap(rw[i][1], rw[j][1]) or _overlap(rw[i][0], rw[j][1])
                if hazard:
                    assert i in anc[j], (model_name, ver, i, j)
        assert max(len(d) for d in deps) <= 6
        names = ["+".join(op.names) if isinstance(op, ConvOp) else "other" for op in ops]
        ms = [names.index("Detect.m.0"), names.index("Detect.m.1")]
        order, cuts = chain_schedule(ops, ms)
        assert sorted(order) == list(range(len(ops))) and cuts[-1] == len(ops)
        pos = {j: k for k, j in enumerate(order)}
        for j, d in enumerate(deps):
            assert all(pos[i] < pos[j] for i in d)
        for m in ms:
            assert pos[m] + 1 in cuts
