"""Live check of the CPU oracle against the UNMODIFIED reference on fresh seeds (not the committed goldens).

Runs only where the reference tree is present (the build container: /root/reference); skipped on the GPU box, where the
committed fixtures under tests/golden/ (tests/test_oracle_golden.py) carry the pin.  TEST INFRASTRUCTURE."""
import numpy as np
import pytest
import torch

from oracle import ref_import
from oracle import yolopoint_oracle as O
from yolopoint_b200.synth import perturb_state_dict, synthetic_frame

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
NAMES = [str(i) for i in range(80)]


@pytest.fixture(scope="module")
def ns():
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ref_import.load()


@pytest.mark.parametrize("model_name,ver", [("YOLOPoint", "n"), ("YOLOPointv52", "n"), ("YOLOPointv52", "s")])
def test_network_forward(ns, model_name, ver):
    """OracleNet == the reference's fused eval forward (src/models/YOLOPoint.py:198-246 / :295-342) on a fresh seed."""
    torch.manual_seed(5)
    m = ns.Model(names=NAMES, version=ver, model_name=model_name)
    sd = perturb_state_dict(m.state_dict(), 5, ver)
    m.load_state_dict(sd)
    m.eval().fuse()
    x = torch.from_numpy(np.random.RandomState(21).rand(2, 3, 96, 128).astype(np.float32))
    with torch.no_grad():
        r = m(x)
    o = O.OracleNet(sd, ver, 80, model_name).forward(x)
    np.testing.assert_allclose(o["semi"].numpy(), r["semi"].numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(o["desc"].numpy(), r["desc"].numpy(), rtol=0, atol=2e-7)
    np.testing.assert_allclose(o["objects"][0].numpy(), r["objects"][0].numpy(), rtol=1e-6, atol=1e-5)


def test_post_processing_functions(ns):
    rs = np.random.RandomState(99)
    # box NMS (src/utils/general_yolo.py:124-235)
    A, nc = 500, 5
    centers = rs.uniform(30, 290, (20, 2))
    xy = centers[rs.randint(0, 20, A)] + rs.normal(0, 5, (A, 2))
    pred = np.concatenate((xy, rs.uniform(15, 70, (A, 2)), rs.uniform(0, 1, (A, 1)) ** 2, rs.uniform(0, 1, (A, nc)) ** 3), -1).astype(np.float32)[None]
    for ct, it, ml, ag in ((0.25, 0.45, False, False), (0.4, 0.45, True, True), (0.3, 0.6, True, False)):
        ref = ns.non_max_suppression(torch.from_numpy(pred.copy()), ct, it, agnostic=ag, multi_label=ml, max_det=300)[0].numpy()
        got = O.non_max_suppression(pred, ct, it, agnostic=ag, multi_label=ml, max_det=300)[0]
        np.testing.assert_array_equal(got, ref)
    # heatmap (src/utils/utils.py:232-262)
    semi = rs.normal(0, 3, (1, 65, 10, 14)).astype(np.float32)
    np.testing.assert_allclose(O.flatten_detection(semi, variant="torch"), ns.flattenDetection(torch.from_numpy(semi)).numpy()[:, 0], rtol=0, atol=1e-7)
    # keypoints (src/utils/utils.py:465-485, 118-182) on a tie-free heatmap
    heat = (((rs.permutation(80 * 112) + 1.0) / (80 * 112 + 1.0)) ** 5).astype(np.float32).reshape(80, 112)
    for thr, r in ((0.02, 4), (0.15, 8)):
        np.testing.assert_array_equal(O.get_pts_from_heatmap(heat, thr, r), ns.getPtsFromHeatmap(heat, thr, r))
    # descriptor sampling (src/evaluations/descriptor_evaluation.py:148-181)
    coarse = rs.normal(0, 1, (1, 32, 10, 14)).astype(np.float32)
    pts = np.stack((rs.randint(0, 112, 150), rs.randint(0, 80, 150), rs.uniform(0, 1, 150))).astype(np.float64)
    np.testing.assert_allclose(O.sample_desc_from_points(coarse, pts), ns.sample_desc_from_points(torch.from_numpy(coarse), pts, "cpu"), rtol=0, atol=2e-7)
    # two-way match (src/demo.py:300-341)
    d1 = rs.normal(0, 1, (32, 200)).astype(np.float32)
    d1 /= np.linalg.norm(d1, axis=0)
    d2 = (d1[:, rs.permutation(200)[:170]] + 0.1 * rs.normal(0, 1, (32, 170))).astype(np.float32)
    d2 /= np.linalg.norm(d2, axis=0)
    np.testing.assert_array_equal(O.nn_match_two_way(d1, d2, 0.7), ns.PointTracker.nn_match_two_way(d1, d2, 0.7))


def test_whole_frame(ns):
    """process_frame == YoloPointFrontend.process_img (src/demo.py:125-230) on a frame seed the goldens do not use."""
    torch.manual_seed(0)
    m = ns.Model(names=NAMES, version="n")
    sd = perturb_state_dict(m.state_dict(), 0, "n")
    m.load_state_dict(sd)
    m.eval().fuse()
    fe = ref_import.make_frontend(ns, m, O.DEFAULT_CFG)
    frame = synthetic_frame(256, 320, 9)
    pts, desc, obj = fe.process_img(frame)
    p2, d2, b2 = O.process_frame(O.OracleNet(sd, "n", 80), frame)
    np.testing.assert_array_equal(p2, pts)
    np.testing.assert_array_equal(b2, obj[0].numpy())
    np.testing.assert_allclose(d2, desc, rtol=0, atol=1e-6)


@pytest.mark.parametrize("shape,crop_resize", [((490, 650), None), ((480, 640), None), ((1216, 1936), [200, 1000, 300, 1900, 1280]),
                                               ((1210, 1930), [100, 900, 50, 1650, 800])])
def test_frontend_host_logic(ns, shape, crop_resize):
    """YoloPointFrontend.preprocess / coordinate restore / template filter (host side of src/demo.py:97-123, 178-228) against the
    reference's own methods on the same frame."""
    import yolopoint_b200 as yp
    rs = np.random.RandomState(shape[0])
    frame = rs.randint(0, 256, shape + (3,)).astype(np.uint8)
    ref = ref_import.make_frontend(ns, None, O.DEFAULT_CFG)
    ref.crop_resize = crop_resize
    ours = yp.YoloPointFrontend(None, {"crop_resize": crop_resize} if crop_resize else None)
    a, b = ref.preprocess(frame), ours.preprocess(frame)
    np.testing.assert_array_equal(a[0], b[0])
    assert tuple(a[1:]) == tuple(b[1:])
    # coordinate restore: the tail of the reference's process_img (src/demo.py:217-228), restated verbatim on copies
    img, cth, ctw, fac = a
    H, W = img.shape[:2]
    pts = np.vstack((rs.randint(4, W - 4, 40), rs.randint(4, H - 4, 40), rs.uniform(0, 1, 40))).astype(np.float64)
    boxes = torch.tensor(rs.uniform(0, 300, (5, 6)).astype(np.float32))
    rp, rb = pts.copy().transpose(), boxes.clone()
    rp[:, 0] = (rp[:, 0] + ctw) / fac
    rp[:, 1] = (rp[:, 1] + cth) / fac
    rp = rp.transpose()
    rb[:, :4] = (rb[:, :4] + torch.tensor([ctw, cth, ctw, cth])) / fac
    if crop_resize:
        rp[:, 0] += crop_resize[2]
        rp[:, 1] += crop_resize[0]
        rb[:, :4] += torch.tensor([crop_resize[2], crop_resize[0], crop_resize[2], crop_resize[0]])
    gp, gb = ours.restore_coords(pts.copy(), boxes.clone(), cth, ctw, fac)
    np.testing.assert_array_equal(gp, rp)
    assert torch.equal(gb, rb)
    # template filter: mask semantics of the reference closure (ones * template == 1)
    template = (rs.rand(H, W) > 0.3).astype(np.float64)
    ours.templates["cam"] = template
    desc = rs.normal(0, 1, (16, 40)).astype(np.float32)
    mask = np.ones((H, W)) * template
    keep = mask[pts[1].astype(int), pts[0].astype(int)] == 1
    fp, fd, dropped = ours.template_filter(pts.copy(), desc, "cam")
    np.testing.assert_array_equal(fp, pts[:, keep])
    np.testing.assert_array_equal(fd, desc[:, keep])
    assert dropped == (not keep.all())
    assert ours.template_filter(pts, desc, "other")[2] is False          # unknown camera: warn and keep everything


@pytest.mark.parametrize("model_name", ["YOLOPoint", "YOLOPointv52"])
def test_fuse_matches_reference(ns, model_name):
    """Model.fuse(): same state-dict keys and folded values as the reference's fuse (src/models/YOLOPoint.py:84-90)."""
    import yolopoint_b200 as yp
    torch.manual_seed(3)
    ref = ns.Model(names=NAMES, version="n", model_name=model_name)
    sd = perturb_state_dict(ref.state_dict(), 3, "n")
    ref.load_state_dict(sd)
    ours = yp.Model(names=NAMES, version="n", model_name=model_name)
    ours.load_state_dict(sd)
    a, b = ref.eval().fuse().state_dict(), ours.eval().fuse().state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        np.testing.assert_allclose(b[k].numpy(), a[k].numpy(), rtol=1e-6, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("seed,maxl", [(1, 2), (2, 4), (3, 6)])
def test_tracker_against_live_reference(ns, seed, maxl, capsys):
    """PointTracker.update / get_tracks against the reference tracker on fresh random sequences and window lengths (the committed
    golden covers one sequence at max_length 4)."""
    import contextlib
    import io
    import yolopoint_b200 as yp
    from oracle.make_golden import tracker_sequence
    ref = ns.PointTracker(max_length=maxl, nn_thresh=0.7)
    ours = yp.PointTracker(max_length=maxl, nn_thresh=0.7)
    for p, d in tracker_sequence(seed=seed, frames=10, n0=90, D=16):
        m = None
        if p is not None:
            prev = ref.last_desc if ref.last_desc is not None else np.zeros((d.shape[0], 0))
            m = O.nn_match_two_way(prev, d, 0.7)
        with contextlib.redirect_stdout(io.StringIO()):
            ref.update(p, d)
        ours.update(p, d, matches=m)
        np.testing.assert_array_equal(ours.tracks, ref.tracks)
        assert ours.track_count == ref.track_count
        for k in (1, 2, maxl):
            np.testing.assert_array_equal(ours.get_tracks(k), ref.get_tracks(k))
    capsys.readouterr()


def test_compat_install_switches_the_reference_seam(ns):
    """yolopoint_b200.compat.install() rebinds the names the reference scripts resolve at run time; the reference's own load_model
    (src/utils/utils.py:55-57) then builds yolopoint_b200.Model with the reference's state dict, for both model families."""
    import yolopoint_b200 as yp
    import yolopoint_b200.compat as compat
    import utils.utils as ru
    import demo
    torch.manual_seed(0)
    ref_sd = ns.Model(names=NAMES, version="n", model_name="YOLOPointv52").state_dict()
    done = compat.install()
    try:
        assert {"models.Model", "utils.utils.nms_fast", "utils.general_yolo.non_max_suppression", "demo.PointTracker"} <= set(done)
        m = ru.load_model(inp_ch=3, names=NAMES, version="n", model_name="YOLOPointv52")        # the call of src/demo.py:45
        assert isinstance(m, yp.Model) and type(m.model).__name__ == "YOLOPointv52"
        m.load_state_dict(ref_sd, strict=True)
        assert demo.PointTracker is yp.PointTracker and ru.getPtsFromHeatmap is yp.getPtsFromHeatmap
        bare = yp.load_model(meta_model=False, model_name="YOLOPoint", width_multiple=0.25, depth_multiple=0.33, inp_ch=3, nc=80,
                             anchors=yp.model.ANCHORS_DEFAULT)
        assert type(bare).__name__ == "YOLOPoint"
    finally:
        compat.uninstall()
    assert ru.nms_fast is ns.nms_fast and demo.PointTracker is ns.PointTracker
    import utils.loss_functions as rl
    ref_infonce = rl.infonce
    done = compat.install(losses=True)
    try:
        assert "utils.loss_functions.infonce" in done and rl.infonce is yp.losses.infonce and ru.getMasks is yp.losses.getMasks
    finally:
        compat.uninstall()
    assert rl.infonce is ref_infonce


def test_tracker_evaluation_variant(ns, capsys):
    """The tracker of src/models/model_wrap.py:410-597 (export_descriptor.py uses it with max_length 2): same tracks, plus get_matches()
    (raw rows for the first pair, then matched coordinates) and clear_desc()."""
    import contextlib
    import io
    import models.model_wrap as mw
    import yolopoint_b200 as yp
    from oracle.make_golden import tracker_sequence
    ref, ours = mw.PointTracker(max_length=2, nn_thresh=0.7), yp.PointTracker(max_length=2, nn_thresh=0.7)
    for f, (p, d) in enumerate(tracker_sequence(seed=7, frames=7, n0=80, D=16)):
        m = None
        if p is not None:
            prev = ref.last_desc if ref.last_desc is not None else np.zeros((d.shape[0], 0))
            m = O.nn_match_two_way(prev, d, 0.7)
        with contextlib.redirect_stdout(io.StringIO()):
            ref.update(p, d)
        ours.update(p, d, matches=m)
        np.testing.assert_array_equal(ours.tracks, ref.tracks)
        if p is not None:
            np.testing.assert_array_equal(ours.get_matches(), ref.get_matches())
        if f == 4:
            ref.clear_desc()
            ours.clear_desc()
    np.testing.assert_array_equal(ours.get_mscores(), ref.get_mscores())
    capsys.readouterr()


def test_edge_cases_match_reference(ns):
    """Empty / degenerate inputs of every post-processing function: the oracle returns what the reference returns (shape, dtype and
    values) and raises where it raises (src/demo.py:300-341, src/utils/utils.py:118-182, 465-485, src/utils/general_yolo.py:124-235,
    src/evaluations/descriptor_evaluation.py:148-181)."""
    rs = np.random.RandomState(3)
    d = rs.normal(0, 1, (32, 10)).astype(np.float32)
    d /= np.linalg.norm(d, axis=0)
    e = np.zeros((32, 0), np.float32)
    match = ns.PointTracker.nn_match_two_way
    for a, b in ((e, d), (d, e), (e, e), (d[:, :1], d[:, :1]), (d, d)):
        r, o = match(a, b, 0.7), O.nn_match_two_way(a, b, 0.7)
        assert r.shape == o.shape and r.dtype == o.dtype
        np.testing.assert_array_equal(o, r)
    for bad in (lambda f: f(d, d, -1.0), lambda f: f(d, d[:16], 0.7)):
        with pytest.raises((ValueError, AssertionError)) as er:
            bad(match)
        with pytest.raises(er.type):
            bad(O.nn_match_two_way)
    # keypoints: nothing above the threshold, one pixel, only border pixels, a constant map (ties: stable order decides)
    h0 = np.zeros((40, 56), np.float32)
    h1 = h0.copy(); h1[20, 30] = 0.9
    hb = h0.copy(); hb[1, 1] = 0.9; hb[38, 54] = 0.8
    hc = np.full((24, 24), 0.5, np.float32)
    for h in (h0, h1, hb, hc):
        r, o = ns.getPtsFromHeatmap(h, 0.1, 4), O.get_pts_from_heatmap(h, 0.1, 4)
        assert np.asarray(r).shape == np.asarray(o).shape
        np.testing.assert_array_equal(np.asarray(o), np.asarray(r))
    for pts in (np.zeros((3, 0)), np.array([[5.0], [6.0], [0.7]])):
        (rp, ri), (op, oi) = ns.nms_fast(pts, 40, 56, 4), O.nms_fast(pts, 40, 56, 4)
        assert rp.shape == op.shape and rp.dtype == op.dtype and ri.dtype == oi.dtype
        np.testing.assert_array_equal(op, rp)
        np.testing.assert_array_equal(oi, ri)
    # box NMS: no candidate in any image of the batch; thresholds outside [0, 1] are refused
    p = np.zeros((2, 100, 6), np.float32)
    r, o = ns.non_max_suppression(torch.from_numpy(p), 0.4, 0.45), O.non_max_suppression(p, 0.4, 0.45)
    assert [tuple(t.shape) for t in r] == [tuple(t.shape) for t in o] == [(0, 6), (0, 6)]
    for f, arg in ((ns.non_max_suppression, torch.from_numpy(p)), (O.non_max_suppression, p)):
        with pytest.raises(AssertionError):
            f(arg, 1.5, 0.45)
        with pytest.raises(AssertionError):
            f(arg, 0.4, -0.1)
    # descriptor sampling with no points
    c = rs.normal(0, 1, (1, 32, 10, 14)).astype(np.float32)
    r, o = ns.sample_desc_from_points(torch.from_numpy(c), np.zeros((3, 0)), "cpu"), O.sample_desc_from_points(c, np.zeros((3, 0)))
    assert r.shape == o.shape == (32, 0) and r.dtype == o.dtype
