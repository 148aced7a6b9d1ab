"""Training losses (yolopoint_b200/losses.py) vs vectors produced by the UNMODIFIED reference (oracle/make_golden.py::golden_losses,
src/utils/loss_functions.py:90-234, 361-481, 600-619).  CPU; deterministic losses bit-close, the sampled descriptor loss statistically."""
import numpy as np
import torch

from yolopoint_b200 import Model, losses as Lz

CFG = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, iou_t=0.2, anchor_t=4.0, label_smoothing=0.0, fl_gamma=0.0)


def test_object_loss_matches_reference(golden):
    g = golden("losses.npz")
    torch.manual_seed(0)
    m = Model(names=[str(i) for i in range(80)], version="n")
    p = [torch.from_numpy(g[f"p{i}"]).requires_grad_(True) for i in range(3)]
    loss, items = Lz.ComputeObjectLoss(m, CFG, "cpu")(p, torch.from_numpy(g["targets"]))
    loss.backward()
    np.testing.assert_allclose(loss.detach().numpy(), g["lobj"], rtol=1e-6)
    np.testing.assert_allclose(items.numpy(), g["lobj_items"], rtol=1e-6)
    np.testing.assert_allclose(p[0].grad.numpy(), g["g0"], rtol=1e-5, atol=1e-9)


def test_detector_loss_and_label_layout_match_reference(golden):
    g = golden("losses.npz")
    lab, mask = torch.from_numpy(g["labels"]), torch.from_numpy(g["mask"])
    np.testing.assert_array_equal(Lz.labels2Dto3D(lab).numpy(), g["labels3d"])
    np.testing.assert_array_equal(Lz.getMasks(mask, "cpu").numpy(), g["mask3d"])
    semi = torch.from_numpy(g["semi"]).requires_grad_(True)
    loss = Lz.ComputeDetectorLoss("cpu")(semi, Lz.labels2Dto3D(lab), Lz.getMasks(mask, "cpu"))
    loss.backward()
    np.testing.assert_allclose(loss.detach().numpy(), g["ldet"], rtol=1e-6)
    np.testing.assert_allclose(semi.grad.numpy(), g["gsemi"], rtol=1e-5, atol=1e-9)


def test_sparse_descriptor_loss_matches_reference_statistically(golden):
    """The loss is a mean over randomly sampled pairs; the reference's numpy stream is not reproduced (losses.py docstring), so the
    mean over repeated draws is compared: reference draws have a std of ~1e-3 on this input."""
    g = golden("losses.npz")
    torch.manual_seed(1)
    d1, d2, mask, Hm = (torch.from_numpy(g[k]) for k in ("d1", "d2", "mask", "Hm"))
    vals = np.array([Lz.descriptor_loss_sparse(d1, d2, mask, Hm, num_samples_per_image=200, num_masked_non_matches_per_match=50).item() for _ in range(8)])
    assert abs(vals.mean() - g["ldesc"].mean()) < 4e-3, (vals, g["ldesc"])
    # identical vs negated descriptors under the identity warp (bilinear sampling at x*(Wc-1)/Wc blends neighbouring cells, as in
    # the reference, so similarities are not exactly +-1)
    eye = torch.eye(3).repeat(d1.shape[0], 1, 1)
    ones = torch.ones_like(mask)
    same = Lz.descriptor_loss_sparse(d1, d1, ones, eye, num_samples_per_image=200, num_masked_non_matches_per_match=50)
    diff = Lz.descriptor_loss_sparse(d1, -d1, ones, eye, num_samples_per_image=200, num_masked_non_matches_per_match=50)
    assert float(diff) > float(same) + 0.5


def test_infonce_matches_reference_statistically(golden):
    """The descriptor loss the reference's training script uses (src/train.py:8 imports ``infonce`` as its descriptor loss;
    src/utils/loss_functions.py:484-597): (1 + K)-way cross entropy of every sampled match against K random non-matches at tau = 0.07.
    Random sampling as in the hinge form, so means over repeated draws are compared (reference std 0.014 on this input)."""
    g, r = golden("losses.npz"), golden("infonce.npz")
    torch.manual_seed(1)
    d1, d2, mask, Hm = (torch.from_numpy(g[k]) for k in ("d1", "d2", "mask", "Hm"))
    vals = np.array([Lz.infonce(d1, d2, mask, Hm, num_samples_per_image=200, num_masked_non_matches_per_match=50).item() for _ in range(16)])
    assert abs(vals.mean() - r["linfonce"].mean()) < 2e-2, (vals.mean(), r["linfonce"].mean())
    eye, ones = torch.eye(3).repeat(d1.shape[0], 1, 1), torch.ones_like(mask)
    same = np.array([Lz.infonce(d1, d1, ones, eye, num_samples_per_image=200, num_masked_non_matches_per_match=50).item() for _ in range(4)])
    assert abs(same.mean() - r["linfonce_same"].mean()) < 2e-2, (same.mean(), r["linfonce_same"].mean())
    # one draw of pairs shared by both forms: the logits are the same similarities the hinge form thresholds; gradients reach both maps
    pairs = Lz.descriptor_pairs(mask, Hm, d1.shape[0], d1.shape[2], d1.shape[3], 200, 50)
    a, b = d1.clone().requires_grad_(True), d2.clone().requires_grad_(True)
    loss = Lz.infonce(a, b, mask, Hm, 200, 50, pairs=pairs)
    loss.backward()
    pos, neg = Lz._pair_similarities(d1, d2, pairs)
    want = -torch.log_softmax(torch.cat((pos[:, None], neg.t()), 1) / 0.07, 1)[:, 0].mean()
    assert abs(float(loss) - float(want)) < 1e-6 and float(a.grad.abs().sum()) > 0 and float(b.grad.abs().sum()) > 0


def test_train_step_options_on_cpu():
    """TrainStep wiring on the CPU (plain PyTorch convolutions): both descriptor losses, the optional gradient clipping of
    src/train.py:249-250 (total gradient norm <= max_norm before Adam), and rejection of an unknown loss name."""
    import pytest
    from yolopoint_b200 import Model
    from yolopoint_b200.trainer import TrainStep, synthetic_sample
    smp = synthetic_sample(2, 64, 96, 3)
    cfg = dict(num_samples_per_image=40, num_masked_non_matches_per_match=10)
    vals = {}
    for name in ("infonce", "hinge"):
        torch.manual_seed(0)
        m = Model(names=[str(i) for i in range(80)], version="n").train()
        ts = TrainStep(m, sparse_cfg=cfg, desc_loss=name, gradclip=0.5)
        torch.manual_seed(5)
        _, parts = ts.losses(smp)
        vals[name] = float(parts["desc"])
        torch.manual_seed(5)
        assert np.isfinite(float(ts.step(smp)))
        assert float(ts.reducer.flat.norm()) <= 0.5 * 1.001
    assert vals["infonce"] > 1.5 and vals["hinge"] < 1.5           # ln(1 + 10) = 2.4 at random init vs a hinge of order 1
    with pytest.raises(ValueError):
        TrainStep(m, desc_loss="triplet")


def test_box_loss_math_matches_autograd(tmp_path):
    """The per-candidate arithmetic of the fused object-loss kernel (yolopoint_b200/csrc/box_loss_math.cuh: box decode, CIoU and its
    gradient carried by dual numbers, BCE-with-logits) compiled for the HOST and compared with PyTorch autograd in fp64 over 20 000
    random candidates: CIoU 1e-6 abs, gradient 2e-6 abs (values of order 1)."""
    import ctypes as C
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = str(tmp_path / "libboxloss_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(here, "box_loss_host.cpp")])
    L = C.CDLL(so)
    rs = np.random.RandomState(0)
    n = 20000
    q = (rs.randn(n, 4) * 2).astype(np.float32)
    an = rs.uniform(0.5, 8, (n, 2)).astype(np.float32)
    tb = np.concatenate([rs.uniform(-0.5, 1.5, (n, 2)), rs.uniform(0.2, 12, (n, 2))], 1).astype(np.float32)
    ci, gr = np.zeros(n, np.float32), np.zeros((n, 4), np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    L.yp_host_candidate_ciou(ptr(q), ptr(an), ptr(tb), n, C.c_float(1e-7), ptr(ci), ptr(gr))
    tq = torch.tensor(q, dtype=torch.float64, requires_grad=True)
    xy, wh = tq[:, 0:2].sigmoid() * 2 - 0.5, (tq[:, 2:4].sigmoid() * 2) ** 2 * torch.tensor(an, dtype=torch.float64)
    c = Lz.ciou_xywh(torch.cat((xy, wh), 1), torch.tensor(tb, dtype=torch.float64))
    c.sum().backward()
    assert np.abs(ci - c.detach().numpy()).max() < 1e-6
    assert np.abs(gr - tq.grad.numpy()).max() < 2e-6
    x, t = (rs.randn(n) * 5).astype(np.float32), rs.rand(n).astype(np.float32)
    lo, dx = np.zeros(n, np.float32), np.zeros(n, np.float32)
    L.yp_host_bce_logits(ptr(x), ptr(t), n, C.c_float(1.7), ptr(lo), ptr(dx))
    tx = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(tx, torch.tensor(t, dtype=torch.float64), pos_weight=torch.tensor([1.7], dtype=torch.float64),
                                                               reduction="none")
    ref.sum().backward()
    np.testing.assert_allclose(lo, ref.detach().numpy(), rtol=2e-6, atol=1e-6)
    assert np.abs(dx - tx.grad.numpy()).max() < 1e-6


def test_in_kernel_target_assignment_matches_build_targets(tmp_path):
    """The target assignment the object-loss kernels evaluate per candidate (plan_candidate of csrc/box_loss_math.cuh, compiled for the
    HOST here) against ComputeObjectLoss.build_targets on 5 000 targets incl. centres exactly on cell and half-cell borders, boxes
    outside the image and extreme aspect ratios: validity, cell, relative box and class bit-identical on all three levels."""
    import ctypes as C
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = str(tmp_path / "libboxloss_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(here, "box_loss_host.cpp")])
    L = C.CDLL(so)
    m = Model(names=[str(i) for i in range(80)], version="n")
    crit = Lz.ComputeObjectLoss(m, CFG, "cpu")
    gen = torch.Generator().manual_seed(0)
    B, nt = 4, 5000
    tg = torch.cat((torch.randint(0, B, (nt, 1), generator=gen).float(), torch.randint(0, 80, (nt, 1), generator=gen).float(),
                    torch.rand(nt, 2, generator=gen) * 1.1 - 0.05, torch.rand(nt, 2, generator=gen) ** 3 * 0.9 + 1e-3), 1)
    tg[:200, 2] = torch.arange(200) / 160.0
    tg[200:400, 3] = torch.arange(200) / 96.0 * 0.5
    grids = ((48, 80), (24, 40), (12, 20))
    plan = crit.build_targets([torch.empty((B, 3, ny, nx, 85), device="meta") for ny, nx in grids], tg)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    t32 = tg.numpy().astype(np.float32).copy()
    for i, lv in enumerate(plan.levels):
        ny, nx = grids[i]
        na, E = 3, 5 * 3 * nt
        an = np.asarray(crit.anchors_host[i], np.float32)
        valid, cell, tbox, cls = np.zeros(E, np.uint8), np.zeros(E, np.int64), np.zeros((E, 4), np.float32), np.zeros(E, np.int32)
        L.yp_host_plan_candidates(ptr(t32), nt, ptr(an), na, nx, ny, C.c_float(CFG["anchor_t"]), ptr(valid), ptr(cell), ptr(tbox), ptr(cls))
        v = lv["valid"].numpy()
        assert v.sum() > 1000
        np.testing.assert_array_equal(valid.astype(bool), v)
        np.testing.assert_array_equal(cell[v], lv["cell"].numpy()[v])
        np.testing.assert_array_equal(tbox[v], lv["tbox"].numpy()[v])
        np.testing.assert_array_equal(cls[v], lv["cls"].numpy()[v])
