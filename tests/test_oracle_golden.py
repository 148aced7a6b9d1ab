"""Pins the CPU oracle (oracle/yolopoint_oracle.py) against vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import yolopoint_oracle as O


def test_box_nms_cases(golden):
    g = golden("box_nms.npz")
    pred = g["pred"]
    for ci, (ct, it, ml, ag, md, cl) in enumerate(g["cases"]):
        classes = None if cl < 0 else [1, 3]
        out = O.non_max_suppression(pred, float(ct), float(it), classes=classes, agnostic=bool(ag),
                                    multi_label=bool(ml), max_det=int(md))
        for b in range(pred.shape[0]):
            ref = g[f"case{ci}_img{b}"]
            assert out[b].shape == ref.shape, (ci, b)
            np.testing.assert_array_equal(out[b], ref)  # bit-exact incl. order


def test_heatmap_variants(golden):
    g = golden("heatmap.npz")
    ht = O.flatten_detection(g["semi"], variant="torch")
    np.testing.assert_allclose(ht, g["heat_torch"][:, 0], rtol=0, atol=1e-7)
    hd = O.flatten_detection(g["semi"][0], variant="demo")
    np.testing.assert_allclose(hd, g["heat_demo0"], rtol=0, atol=1e-7)


def test_keypoints(golden):
    g = golden("keypoints.npz")
    for ci, (thr, r) in enumerate(g["kcases"]):
        pts = O.get_pts_from_heatmap(g["heat"], float(thr), int(r))
        np.testing.assert_array_equal(pts, g[f"pts{ci}"])
    np.testing.assert_array_equal(O.get_pts_from_heatmap(g["border_heat"], 0.1, 4), g["border_pts"])
    single = np.zeros((32, 48), np.float32); single[12, 17] = .7
    np.testing.assert_array_equal(O.get_pts_from_heatmap(single, 0.1, 4), g["single_pts"])


def test_filter_points(golden):
    g = golden("filter_pts.npz")
    H, W = g["HW"]
    np.testing.assert_array_equal(O.filter_points_in_boxes(g["pts"], g["boxes"], int(H), int(W)), g["out"])


def test_sample_desc(golden):
    g = golden("sample_desc.npz")
    d = O.sample_desc_from_points(g["coarse"], g["pts"])
    assert d.shape == g["desc"].shape
    np.testing.assert_allclose(d, g["desc"], rtol=0, atol=2e-7)


def test_match(golden):
    g = golden("match.npz")
    np.testing.assert_array_equal(O.nn_match_two_way(g["desc1"], g["desc2"], 0.7), g["m07"])
    np.testing.assert_array_equal(O.nn_match_two_way(g["desc1"], g["desc2"], 0.3), g["m03"])
    assert O.nn_match_two_way(g["desc1"][:, :0], g["desc2"], 0.7).shape == (3, 0)
    try:
        O.nn_match_two_way(g["desc1"], g["desc2"], -1.0)
        assert False
    except ValueError:
        pass


def _model_sd(version):
    from yolopoint_b200 import Model
    from yolopoint_b200.synth import perturb_state_dict
    torch.manual_seed(0)
    m = Model(names=[str(i) for i in range(80)], version=version)
    return m, perturb_state_dict(m.state_dict(), 0, version)


def test_network_forward_n(golden):
    g = golden("net_n_64x96.npz")
    _, sd = _model_sd("n")
    net = O.OracleNet(sd, "n", 80)
    o = net.forward(torch.from_numpy(g["x"]))
    np.testing.assert_allclose(o["semi"].numpy(), g["semi"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(o["desc"].numpy(), g["desc"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(o["objects"][0].numpy(), g["pred"], rtol=1e-5, atol=1e-4)
    for i in range(3):
        np.testing.assert_allclose(o["objects"][1][i].numpy(), g[f"raw{i}"], rtol=0, atol=2e-4)


def test_e2e_n(golden):
    from yolopoint_b200.synth import synthetic_frame
    g = golden("e2e_n_480x640.npz")
    _, sd = _model_sd("n")
    net = O.OracleNet(sd, "n", 80)
    res = [O.process_frame(net, synthetic_frame(480, 640, s)) for s in (0, 1)]
    for i, (pts, desc, boxes) in enumerate(res):
        np.testing.assert_array_equal(pts[:2], g[f"pts{i}"][:2])  # keypoint coordinates bit-exact
        np.testing.assert_allclose(pts[2], g[f"pts{i}"][2], rtol=0, atol=1e-6)
        np.testing.assert_allclose(desc, g[f"desc{i}"], rtol=0, atol=1e-5)
        assert boxes.shape == g[f"boxes{i}"].shape
        np.testing.assert_allclose(boxes, g[f"boxes{i}"], rtol=0, atol=1e-3)
    m = O.nn_match_two_way(res[0][1], res[1][1], 0.7)
    np.testing.assert_array_equal(m[:2], g["matches"][:2])


def test_homography_adaptation_contract(golden):
    """warp_image_batch (bilinear / nearest) and the heatmap aggregation of the homography-adaptation export (SURVEY.md section 8f
    rank 2) against the reference's own functions: the contract the next kernel is built to."""
    g = golden("homography.npz")
    np.testing.assert_allclose(O.warp_image_batch(g["heat"], g["inv_homographies"]), g["warp_bilinear"], rtol=0, atol=5e-7)
    wn = O.warp_image_batch(g["heat"], g["inv_homographies"], mode="nearest")
    assert (wn != g["warp_nearest"]).mean() < 1e-3            # a source coordinate within 1 ulp of x.5 may round the other way
    # identity homography: linspace coordinates are not exact integers in fp32, so the warp is the identity only to ~1e-5 (reference too)
    np.testing.assert_allclose(O.warp_image_batch(g["heat"][:1], g["inv_homographies"][:1]), g["heat"][:1], rtol=0, atol=2e-5)
    agg = O.homography_adaptation(g["heat"], g["mask"], g["inv_homographies"])
    assert np.array_equal(np.isnan(agg), np.isnan(g["aggregated"]))
    np.testing.assert_allclose(np.nan_to_num(agg), np.nan_to_num(g["aggregated"]), rtol=0, atol=1e-6)
