"""GPU parity of the post-processing kernels through the C ABI / reference-named API:
bit-exact indices against the golden vectors (produced by the unmodified reference) and against the CPU oracle on
seeded random inputs; floats within the tolerance written next to each check."""
import numpy as np
import pytest
import torch

import yolopoint_b200 as yp
from oracle import yolopoint_oracle as O
from yolopoint_b200 import ops

pytestmark = pytest.mark.gpu


def test_box_nms_golden(golden):
    g = golden("box_nms.npz")
    pred = torch.from_numpy(g["pred"]).cuda()
    for ci, (ct, it, ml, ag, md, cl) in enumerate(g["cases"]):
        classes = None if cl < 0 else [1, 3]
        out = yp.non_max_suppression(pred, float(ct), float(it), classes=classes, agnostic=bool(ag), multi_label=bool(ml), max_det=int(md))
        for b in range(pred.shape[0]):
            ref = g[f"case{ci}_img{b}"]
            got = out[b].cpu().numpy()
            assert got.shape == ref.shape, (ci, b, got.shape, ref.shape)
            np.testing.assert_array_equal(got, ref)   # identical boxes in identical order: bit-exact


@pytest.mark.parametrize("A,nc,seed", [(25200, 80, 0), (3000, 1, 1), (18900, 7, 2)])
def test_box_nms_vs_oracle_random(A, nc, seed):
    from oracle.make_golden import clustered_pred
    pred = clustered_pred(np.random.RandomState(seed), 2, A, nc, size=640.0)
    pred[..., 4] *= (np.random.RandomState(seed + 9).uniform(0, 1, pred.shape[:2]) < 0.05)   # ~5 % objectness survivors
    for kw in (dict(conf_thres=0.4, iou_thres=0.45, multi_label=True, agnostic=True, max_det=1000),
               dict(conf_thres=0.25, iou_thres=0.45, multi_label=False, agnostic=False, max_det=300)):
        ref = O.non_max_suppression(pred, **kw)
        got = yp.non_max_suppression(torch.from_numpy(pred).cuda(), **kw)
        for b in range(2):
            assert got[b].shape == ref[b].shape, (kw, b, got[b].shape, ref[b].shape)
            np.testing.assert_array_equal(got[b].cpu().numpy(), ref[b])


def test_box_nms_empty_and_errors():
    pred = torch.zeros(2, 100, 85).cuda()
    out = yp.non_max_suppression(pred, 0.25, 0.45)
    assert all(o.shape == (0, 6) for o in out)
    with pytest.raises(AssertionError):
        yp.non_max_suppression(pred, 1.5, 0.45)


def test_heatmap_golden(golden):
    g = golden("heatmap.npz")
    semi = torch.from_numpy(g["semi"]).cuda()
    ht = yp.flattenDetection(semi)
    assert ht.shape == (2, 1, 96, 128)
    np.testing.assert_allclose(ht.cpu().numpy(), g["heat_torch"], rtol=0, atol=1e-6)     # fp32 softmax: 1e-6 abs
    hd = ops.heatmap(semi[:1].contiguous(), "nchw", variant=1)
    np.testing.assert_allclose(hd[0].cpu().numpy(), g["heat_demo0"], rtol=0, atol=1e-6)
    nhwc = semi.permute(0, 2, 3, 1).contiguous()
    hn = ops.heatmap(nhwc, "nhwc", variant=0)
    assert torch.equal(hn, ht[:, 0])                                                       # layout independent, bit-exact


def test_keypoints_golden(golden):
    g = golden("keypoints.npz")
    for ci, (thr, r) in enumerate(g["kcases"]):
        pts = yp.getPtsFromHeatmap(g["heat"], float(thr), int(r))
        np.testing.assert_array_equal(pts, g[f"pts{ci}"])                                   # bit-exact incl. order
    np.testing.assert_array_equal(yp.getPtsFromHeatmap(g["border_heat"], 0.1, 4), g["border_pts"])
    single = np.zeros((32, 48), np.float32); single[12, 17] = .7
    np.testing.assert_array_equal(yp.getPtsFromHeatmap(single, 0.1, 4), g["single_pts"])
    assert yp.getPtsFromHeatmap(np.zeros((32, 48), np.float32), 0.1, 4).shape == (3, 0)


@pytest.mark.parametrize("H,W,r,thr,seed", [(480, 640, 4, 0.015, 0), (640, 640, 8, 0.12, 1), (736, 1280, 4, 0.085, 2), (96, 96, 2, 0.3, 3)])
def test_keypoints_vs_oracle_random(H, W, r, thr, seed):
    rs = np.random.RandomState(seed)
    heat = (((rs.permutation(H * W) + 1.0) / (H * W + 1.0)) ** 8).astype(np.float32).reshape(H, W)   # tie-free
    ref = O.get_pts_from_heatmap(heat, thr, r)
    got = yp.getPtsFromHeatmap(heat, thr, r)
    np.testing.assert_array_equal(got, ref)


def test_keypoints_worst_case_chain():
    """A monotone ramp forces one decision per round in the parallel fixed point; result must still be exact."""
    H, W = 32, 640
    heat = np.zeros((H, W), np.float32)
    heat[16, :] = np.linspace(0.2, 0.9, W, dtype=np.float32)
    np.testing.assert_array_equal(yp.getPtsFromHeatmap(heat, 0.1, 4), O.get_pts_from_heatmap(heat, 0.1, 4))


def test_filter_points_golden(golden):
    g = golden("filter_pts.npz")
    H, W = (int(v) for v in g["HW"])
    # rebuild a heatmap whose NMS survivors are exactly g["pts"]: isolated peaks (nms_dist 0 keeps everything)
    heat = np.zeros((H, W), np.float32)
    p = g["pts"]
    heat[p[1].astype(int), p[0].astype(int)] = p[2].astype(np.float32)
    boxes = torch.from_numpy(g["boxes"]).cuda().view(1, -1, 6).contiguous()
    cnt = torch.tensor([boxes.shape[1]], dtype=torch.int32).cuda()
    pts, count = ops.keypoints(torch.from_numpy(heat).cuda()[None], 0.015, 0, 0, boxes, cnt, max_pts=4096)
    got = pts[0, :int(count.item())].double().cpu().numpy().T
    ref = g["out"]
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got, ref)


def test_sample_desc_golden(golden):
    g = golden("sample_desc.npz")
    d = yp.sample_desc_from_points(torch.from_numpy(g["coarse"]), g["pts"], "cuda")
    assert d.shape == g["desc"].shape
    np.testing.assert_allclose(d, g["desc"], rtol=0, atol=1e-6)                             # fp32 bilinear + norm: 1e-6 abs
    assert yp.sample_desc_from_points(torch.from_numpy(g["coarse"]), np.zeros((3, 0)), "cuda").shape == (64, 0)


def test_match_golden(golden):
    g = golden("match.npz")
    m = yp.nn_match_two_way(g["desc1"], g["desc2"], 0.7)
    np.testing.assert_array_equal(m[:2], g["m07"][:2])                                      # indices bit-exact
    np.testing.assert_allclose(m[2], g["m07"][2], rtol=0, atol=1e-4)                        # scores: 1e-4 abs (sqrt of fp32 dot)
    assert yp.nn_match_two_way(g["desc1"], g["desc2"], 0.3).shape == g["m03"].shape
    assert yp.nn_match_two_way(g["desc1"][:, :0], g["desc2"], 0.7).shape == (3, 0)
    with pytest.raises(ValueError):
        yp.nn_match_two_way(g["desc1"], g["desc2"], -1.0)


@pytest.mark.parametrize("N1,N2,D", [(512, 512, 64), (1000, 1300, 128), (2048, 2048, 256), (4096, 4000, 192)])
def test_match_vs_oracle_planted(N1, N2, D):
    rs = np.random.RandomState(N1 + D)
    d1 = rs.normal(0, 1, (D, N1)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=0)
    perm = rs.permutation(N1)[:min(N1, N2)]
    d2 = rs.normal(0, 1, (D, N2)).astype(np.float32)
    d2[:, :len(perm)] = d1[:, perm] + 0.05 * rs.normal(0, 1, (D, len(perm))).astype(np.float32)
    d2 /= np.linalg.norm(d2, axis=0)
    ref = O.nn_match_two_way(d1, d2, 0.7)
    got = yp.nn_match_two_way(d1, d2, 0.7)
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got[:2], ref[:2])
    np.testing.assert_allclose(got[2], ref[2], rtol=0, atol=1e-4)


@pytest.mark.parametrize("N1,N2,D", [(700, 900, 64), (2048, 2300, 128), (4096, 4000, 256)])
def test_match_tensor_core_path_vs_oracle(N1, N2, D):
    """The tcgen05 row-minimum passes (3xTF32) give the same matches as the oracle: indices bit-exact, scores 1e-4 abs; device-side
    counts mask the padded rows / columns."""
    rs = np.random.RandomState(N2 + D)
    d1 = rs.normal(0, 1, (D, N1)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=0)
    perm = rs.permutation(N1)[:min(N1, N2)]
    d2 = rs.normal(0, 1, (D, N2)).astype(np.float32)
    d2[:, :len(perm)] = d1[:, perm] + 0.05 * rs.normal(0, 1, (D, len(perm))).astype(np.float32)
    d2 /= np.linalg.norm(d2, axis=0)
    ref = O.nn_match_two_way(d1, d2, 0.7)
    a = torch.from_numpy(np.ascontiguousarray(d1.T)).cuda()
    b = torch.from_numpy(np.ascontiguousarray(d2.T)).cuda()
    m, cnt = ops.match_two_way(a, None, b, None, 0.7, algo="tc")
    got = m[:int(cnt.item())].cpu().numpy().T.astype(np.float64)
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got[:2], ref[:2])
    np.testing.assert_allclose(got[2], ref[2], rtol=0, atol=1e-4)
    # capacity > count: rows / columns beyond the device-side counts must be ignored
    a2 = torch.cat((a, torch.randn(300, D, device="cuda")), 0).contiguous()
    b2 = torch.cat((b, torch.randn(500, D, device="cuda")), 0).contiguous()
    n1 = torch.tensor([N1], dtype=torch.int32, device="cuda"); n2 = torch.tensor([N2], dtype=torch.int32, device="cuda")
    m2, cnt2 = ops.match_two_way(a2, n1, b2, n2, 0.7, algo="tc")
    assert int(cnt2.item()) == int(cnt.item()) and torch.equal(m2[:int(cnt2.item())], m[:int(cnt.item())])


def test_match_full_size_properties():
    """BASELINE config 4 upper size (16384 x 16384, D=256): too slow for the CPU oracle, so check size-independent
    properties: matching a set against a permutation of itself returns the permutation with distance ~0, symmetric."""
    N, D = 16384, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    d1 = torch.randn(N, D, generator=g, device="cuda"); d1 /= d1.norm(dim=1, keepdim=True)
    perm = torch.randperm(N, generator=g, device="cuda")
    d2 = d1[perm].contiguous()
    m, cnt = ops.match_two_way(d1, None, d2, None, 0.7)
    assert int(cnt.item()) == N
    inv = torch.empty_like(perm); inv[perm] = torch.arange(N, device="cuda")
    assert torch.equal(m[:, 0].long(), torch.arange(N, device="cuda")) and torch.equal(m[:, 1].long(), inv)
    assert float(m[:, 2].max()) < 3e-3   # sqrt(2-2*d) with d = 1 - O(1e-7): self-distance noise is O(1e-3)
    m2, cnt2 = ops.match_two_way(d2, None, d1, None, 0.7)
    assert int(cnt2.item()) == N and torch.equal(m2[:, 1].long(), perm)


def test_point_tracker_with_cuda_match_equals_reference(golden, capsys):
    """PointTracker.update with its own match (the CUDA two-way match instead of the reference's numpy one) reproduces the track
    matrices the reference's tracker produced over the synthetic sequence (src/demo.py:358-441), bit for bit in the ids; the running
    mean score column agrees to the match-score tolerance (1e-4 abs)."""
    g = golden("tracker.npz")
    none = set(int(i) for i in g["none_frames"])
    trk = yp.PointTracker(max_length=4, nn_thresh=0.7)
    f = 0
    while f"tracks{f}" in g.files:
        if f in none:
            trk.update(None, None)
        else:
            trk.update(g[f"pts{f}"], g[f"desc{f}"])
        ref = g[f"tracks{f}"]
        assert trk.tracks.shape == ref.shape, f
        np.testing.assert_array_equal(trk.tracks[:, 0], ref[:, 0])
        np.testing.assert_array_equal(trk.tracks[:, 2:], ref[:, 2:])
        np.testing.assert_allclose(trk.tracks[:, 1], ref[:, 1], rtol=0, atol=1e-4)
        assert trk.track_count == int(g[f"count{f}"])
        f += 1
    capsys.readouterr()


def test_warp_image_batch_and_homography_adaptation_golden(golden):
    """SURVEY.md section 8f rank 2 against what the unmodified reference produced (tests/golden/homography.npz): the bilinear warp to
    1e-6 abs, the nearest warp identical except where a source coordinate sits within an ulp of x.5, the fused aggregation
    (product, two warps, two sums, division in one kernel) to 1e-6 abs with the same NaN pattern."""
    g = golden("homography.npz")
    heat, mask, hinv = (torch.from_numpy(g[k]).cuda() for k in ("heat", "mask", "inv_homographies"))
    wb = yp.warp_image_batch(heat, hinv, mode="bilinear")
    np.testing.assert_allclose(wb.cpu().numpy(), g["warp_bilinear"], rtol=0, atol=1e-6)
    wn = yp.warp_image_batch(heat, hinv, mode="nearest").cpu().numpy()
    assert (wn != g["warp_nearest"]).mean() < 1e-3
    agg = yp.homography_adaptation(heat, mask, hinv)[0].cpu().numpy()
    assert np.array_equal(np.isnan(agg), np.isnan(g["aggregated"]))
    np.testing.assert_allclose(np.nan_to_num(agg), np.nan_to_num(g["aggregated"]), rtol=0, atol=1e-6)
    # host input -> host output, 2-D / RGB forms of the reference's signature
    one = yp.warp_image_batch(g["heat"][0, 0], g["inv_homographies"][0])
    assert not one.is_cuda and one.shape == (1, 1, 48, 64)
    np.testing.assert_allclose(one.numpy()[0, 0], g["warp_bilinear"][0, 0], rtol=0, atol=1e-6)


@pytest.mark.parametrize("B,H,W", [(100, 120, 160), (7, 33, 57)])
def test_homography_adaptation_vs_oracle_random(B, H, W):
    """The export's real shape class (100 warped copies per image): random perspective homographies, random masks with holes."""
    rs = np.random.RandomState(B + H)
    heat = rs.rand(B, 1, H, W).astype(np.float32)
    mask = (rs.rand(B, 1, H, W) > 0.2).astype(np.float32)
    hinv = np.tile(np.eye(3, dtype=np.float32), (B, 1, 1))
    hinv[:, :2, :2] += rs.uniform(-0.25, 0.25, (B, 2, 2)).astype(np.float32)
    hinv[:, :2, 2] += rs.uniform(-0.3, 0.3, (B, 2)).astype(np.float32)
    hinv[:, 2, :2] += rs.uniform(-0.15, 0.15, (B, 2)).astype(np.float32)
    ref = O.homography_adaptation(heat, mask, hinv)
    got = yp.homography_adaptation(torch.from_numpy(heat).cuda(), torch.from_numpy(mask).cuda(), torch.from_numpy(hinv).cuda())[0].cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(np.nan_to_num(got), np.nan_to_num(ref), rtol=0, atol=2e-6)
    wref = O.warp_image_batch(heat[:5], hinv[:5])
    wgot = yp.warp_image_batch(torch.from_numpy(heat[:5]).cuda(), torch.from_numpy(hinv[:5]).cuda()).cpu().numpy()
    np.testing.assert_allclose(wgot, wref, rtol=0, atol=2e-6)
    # letterbox slicing as the reference applies it
    padded = yp.homography_adaptation(torch.from_numpy(heat).cuda(), torch.from_numpy(mask).cuda(), torch.from_numpy(hinv).cuda(), pad=(4, 6, 3, 5))
    assert padded.shape == (1, len(range(H)[4:W - 6]), len(range(W)[3:H - 5]))       # (:103-109 bound the rows by the WIDTH and vice versa)


@pytest.mark.parametrize("cap,algo", [(4096, "tc"), (1024, "simt")])
def test_two_way_matcher_graph_equals_direct_call(cap, algo):
    """ops.TwoWayMatcher (static operands + one CUDA graph per call) against ops.match_two_way on successive descriptor sets of
    different sizes below the capacity: identical matches and counts (the device-side counts are read at replay time)."""
    D = 256
    matcher = ops.TwoWayMatcher(cap, cap, D, "cuda", 0.7, algo=algo)
    g = torch.Generator(device="cpu").manual_seed(cap)
    for n1, n2 in ((cap, cap), (cap - 37, cap - 500), (cap // 2, cap // 3)):
        d1 = torch.randn(n1, D, generator=g); d1 /= d1.norm(dim=1, keepdim=True)
        d2 = torch.randn(n2, D, generator=g)
        k = min(n1, n2)
        d2[:k] = d1[torch.randperm(n1, generator=g)[:k]] + 0.05 * torch.randn(k, D, generator=g)
        d2 /= d2.norm(dim=1, keepdim=True)
        d1, d2 = d1.cuda().contiguous(), d2.cuda().contiguous()
        m_ref, c_ref = ops.match_two_way(d1, None, d2, None, 0.7, algo=algo)
        m, c = matcher(d1, d2)
        torch.cuda.synchronize()
        n = int(c_ref.item())
        assert int(c.item()) == n and n > 0
        assert torch.equal(m[:n], m_ref[:n])
