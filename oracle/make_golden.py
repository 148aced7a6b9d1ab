"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE.  Run once in the build container (``python -m oracle.make_golden``); the GPU
box has no /root/reference, so the fixtures are committed.  Every array below is produced by the
reference's own functions (file:line cited per block); inputs are stored next to the outputs or are
re-derivable from the seeds recorded in the file.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle import yolopoint_oracle as O  # noqa: E402
from yolopoint_b200.synth import perturb_state_dict, synthetic_frame  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
NAMES80 = [str(i) for i in range(80)]


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB", {k: getattr(v, 'shape', None) for k, v in arrs.items()})


def clustered_pred(rs, B, A, nc, size=320.0):
    """Synthetic Detect output rows (xywh, obj, cls) with overlapping clusters so NMS has work to do."""
    centers = rs.uniform(40, size - 40, (B, 24, 2))
    which = rs.randint(0, 24, (B, A))
    xy = np.take_along_axis(centers, which[..., None].repeat(2, -1), 1) + rs.normal(0, 6, (B, A, 2))
    wh = rs.uniform(20, 80, (B, A, 2))
    obj = rs.uniform(0, 1, (B, A, 1)) ** 2
    cls = rs.uniform(0, 1, (B, A, nc)) ** 3
    return np.concatenate((xy, wh, obj, cls), -1).astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_import.load()
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- 1. seeded initial state dicts (src/models/YOLOPoint.py:17-100) --------------------
    for ver in ("n", "s"):
        torch.manual_seed(0)
        m = ns.Model(names=NAMES80, version=ver)
        sd = m.state_dict()
        keys = np.array(list(sd.keys()))
        sums = np.array([float(v.double().sum()) for v in sd.values()])
        asums = np.array([float(v.double().abs().sum()) for v in sd.values()])
        shapes = np.array([str(tuple(v.shape)) for v in sd.values()])
        pnames = np.array([n for n, _ in m.named_parameters()])
        save(f"state_{ver}.npz", keys=keys, sums=sums, asums=asums, shapes=shapes, param_names=pnames,
             stride=m.model.Detect.stride.numpy(), anchors=m.model.Detect.anchors.numpy())

    # ---- 2. network forward, N model, B=2, 64x96 (src/models/YOLOPoint.py:198-246, fused) ---
    torch.manual_seed(0)
    m = ns.Model(names=NAMES80, version="n")
    sd = perturb_state_dict(m.state_dict(), 0, "n")
    m.load_state_dict(sd); m.eval().fuse()
    x = torch.from_numpy(np.random.RandomState(7).rand(2, 3, 64, 96).astype(np.float32))
    with torch.no_grad():
        o = m(x)
    save("net_n_64x96.npz", x=x.numpy(), semi=o["semi"].numpy(), desc=o["desc"].numpy(),
         pred=o["objects"][0].numpy(), raw0=o["objects"][1][0].numpy(), raw1=o["objects"][1][1].numpy(),
         raw2=o["objects"][1][2].numpy())

    # ---- 3. whole-frame pipeline (src/demo.py:125-230 + 300-341) -----------------------------
    for ver, (H, W) in (("n", (480, 640)), ("s", (640, 640))):
        torch.manual_seed(0)
        m = ns.Model(names=NAMES80, version=ver)
        sd = perturb_state_dict(m.state_dict(), 0, ver)
        m.load_state_dict(sd); m.eval().fuse()
        fe = ref_import.make_frontend(ns, m, O.DEFAULT_CFG)
        res = []
        for seed in (0, 1):
            pts, desc, obj = fe.process_img(synthetic_frame(H, W, seed))
            res.append((pts, desc, obj[0].numpy()))
        matches = ns.PointTracker.nn_match_two_way(res[0][1], res[1][1], O.DEFAULT_CFG["nn_thresh"])
        save(f"e2e_{ver}_{H}x{W}.npz", pts0=res[0][0], desc0=res[0][1].astype(np.float32), boxes0=res[0][2],
             pts1=res[1][0], desc1=res[1][1].astype(np.float32), boxes1=res[1][2], matches=matches)

    golden_v52(ns)

    # ---- 4. box NMS (src/utils/general_yolo.py:124-235) ---------------------------------------
    rs = np.random.RandomState(11)
    pred = clustered_pred(rs, 2, 600, 7)
    arrs = dict(pred=pred)
    cases = [(0.25, 0.45, False, False, 300, None), (0.4, 0.45, True, True, 1000, None),
             (0.3, 0.6, True, False, 20, None), (0.25, 0.45, True, False, 300, [1, 3])]
    for ci, (ct, it, ml, ag, md, cl) in enumerate(cases):
        out = ns.non_max_suppression(torch.from_numpy(pred.copy()), ct, it, classes=cl, agnostic=ag,
                                     multi_label=ml, max_det=md)
        for b, t in enumerate(out):
            arrs[f"case{ci}_img{b}"] = t.numpy()
    arrs["cases"] = np.array([[c[0], c[1], float(c[2]), float(c[3]), c[4], -1 if c[5] is None else 13] for c in cases])
    save("box_nms.npz", **arrs)

    # ---- 5. heatmap (src/utils/utils.py:232-262 ; src/demo.py:140-150) -------------------------
    semi = (rs.normal(0, 3, (2, 65, 12, 16))).astype(np.float32)
    heat_t = ns.flattenDetection(torch.from_numpy(semi)).numpy()
    dense = np.exp(semi[0]); dense = dense / (np.sum(dense, axis=0) + .00001)
    nodust = dense[:-1].transpose(1, 2, 0)
    heat_d = np.transpose(np.reshape(nodust, [12, 16, 8, 8]), [0, 2, 1, 3]).reshape(96, 128)
    save("heatmap.npz", semi=semi, heat_torch=heat_t, heat_demo0=heat_d)

    # ---- 6. keypoints (src/utils/utils.py:465-485, 118-182) -------------------------------------
    heat = (((rs.permutation(96 * 128) + 1.0) / (96 * 128 + 1.0)) ** 6).astype(np.float32).reshape(96, 128)
    cand = heat[heat >= 0.015]
    assert len(np.unique(cand)) == cand.size, "golden heatmap must be tie-free (the reference's sorts are unstable on ties)"
    arrs = dict(heat=heat)
    kcases = [(0.015, 4), (0.12, 8), (0.3, 2), (0.999999, 4)]
    for ci, (thr, r) in enumerate(kcases):
        arrs[f"pts{ci}"] = ns.getPtsFromHeatmap(heat, thr, r)
    border = np.zeros((32, 48), np.float32); border[10, 2] = .9; border[10, 5] = .8; border[20, 30] = .5
    arrs["border_heat"] = border
    arrs["border_pts"] = ns.getPtsFromHeatmap(border, 0.1, 4)
    single = np.zeros((32, 48), np.float32); single[12, 17] = .7
    arrs["single_pts"] = ns.getPtsFromHeatmap(single, 0.1, 4)
    arrs["kcases"] = np.array(kcases)
    save("keypoints.npz", **arrs)

    # ---- 7. keypoint-in-box filter (src/demo.py:178-198) ----------------------------------------
    H, W = 96, 128
    pts = ns.getPtsFromHeatmap(heat, 0.015, 2)
    boxes = np.array([[10.4, 8.6, 50.5, 40.5, .9, 1], [-3.2, 60.1, 30.7, 90.9, .8, 0], [100.5, -5.0, 140.2, 30.0, .7, 2],
                      [60.0, 50.0, 60.4, 80.0, .6, 3], [70.5, 70.5, 90.5, 200.0, .5, 1]], np.float32)
    fe.filter_pts = True

    def ref_filter(obj_preds, points, im_shape):  # verbatim semantics of the closure at src/demo.py:179-198
        mask = np.ones(im_shape)
        points = points.transpose()
        points2 = points[:, :2].astype(int)
        for *xyxy, _, _ in obj_preds:
            x0, y0, x1, y1 = np.rint(xyxy).astype(int)
            mask[y0:y1, x0:x1] = 0
        return points[mask[points2[:, 1], points2[:, 0]] == 1].transpose()

    save("filter_pts.npz", pts=pts, boxes=boxes, out=ref_filter(boxes, pts, (H, W)), HW=np.array([H, W]))

    # ---- 8. descriptor sampling (src/evaluations/descriptor_evaluation.py:148-181) ---------------
    coarse = rs.normal(0, 1, (1, 64, 12, 16)).astype(np.float32)
    coarse /= np.linalg.norm(coarse, axis=1, keepdims=True)
    spts = np.stack((rs.randint(0, 128, 200), rs.randint(0, 96, 200), rs.uniform(0, 1, 200))).astype(np.float64)
    sdesc = ns.sample_desc_from_points(torch.from_numpy(coarse), spts, "cpu")
    save("sample_desc.npz", coarse=coarse, pts=spts, desc=sdesc)

    # ---- 9. two-way match (src/demo.py:300-341) ---------------------------------------------------
    d1 = rs.normal(0, 1, (64, 300)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=0)
    perm = rs.permutation(300)[:260]
    d2 = d1[:, perm] + 0.08 * rs.normal(0, 1, (64, 260)).astype(np.float32); d2 /= np.linalg.norm(d2, axis=0)
    d2[:, 200:] = rs.normal(0, 1, (64, 60)); d2 /= np.linalg.norm(d2, axis=0)
    d2 = d2.astype(np.float32)
    save("match.npz", desc1=d1, desc2=d2, m07=ns.PointTracker.nn_match_two_way(d1, d2, 0.7),
         m03=ns.PointTracker.nn_match_two_way(d1, d2, 0.3))


def golden_v52(ns=None):
    """YOLOPointv52 (src/models/YOLOPoint.py:248-342): seeded state dict, network forward and the
    whole-frame pipeline, all produced by the reference's own Model(model_name='YOLOPointv52') / YoloPointFrontend.process_img."""
    ns = ns or ref_import.load()
    name = "YOLOPointv52"
    for ver in ("n", "s"):
        torch.manual_seed(0)
        m = ns.Model(names=NAMES80, version=ver, model_name=name)
        sd = m.state_dict()
        save(f"state_v52{ver}.npz", keys=np.array(list(sd.keys())), sums=np.array([float(v.double().sum()) for v in sd.values()]),
             asums=np.array([float(v.double().abs().sum()) for v in sd.values()]), shapes=np.array([str(tuple(v.shape)) for v in sd.values()]),
             param_names=np.array([n for n, _ in m.named_parameters()]), stride=m.model.Detect.stride.numpy(), anchors=m.model.Detect.anchors.numpy())
    torch.manual_seed(0)
    m = ns.Model(names=NAMES80, version="n", model_name=name)
    m.load_state_dict(perturb_state_dict(m.state_dict(), 0, "n")); m.eval().fuse()
    x = torch.from_numpy(np.random.RandomState(7).rand(2, 3, 64, 96).astype(np.float32))
    with torch.no_grad():
        o = m(x)
    save("net_v52n_64x96.npz", x=x.numpy(), semi=o["semi"].numpy(), desc=o["desc"].numpy(), pred=o["objects"][0].numpy(),
         raw0=o["objects"][1][0].numpy(), raw1=o["objects"][1][1].numpy(), raw2=o["objects"][1][2].numpy())
    ver, (H, W) = "s", (640, 640)
    torch.manual_seed(0)
    m = ns.Model(names=NAMES80, version=ver, model_name=name)
    m.load_state_dict(perturb_state_dict(m.state_dict(), 0, ver)); m.eval().fuse()
    fe = ref_import.make_frontend(ns, m, O.DEFAULT_CFG)
    res = []
    for seed in (0, 1):
        pts, desc, obj = fe.process_img(synthetic_frame(H, W, seed))
        res.append((pts, desc, obj[0].numpy()))
    matches = ns.PointTracker.nn_match_two_way(res[0][1], res[1][1], O.DEFAULT_CFG["nn_thresh"])
    save(f"e2e_v52{ver}_{H}x{W}.npz", pts0=res[0][0], desc0=res[0][1].astype(np.float32), boxes0=res[0][2],
         pts1=res[1][0], desc1=res[1][1].astype(np.float32), boxes1=res[1][2], matches=matches)


def tracker_sequence(seed=0, frames=8, n0=160, D=32):
    """Synthetic observation sequence for the tracker: a pool of unit descriptors with drifting positions; every frame sees a random
    subset (plus a few new points), descriptors jittered; frame 3 is empty and frame 5 is a dropped frame (None)."""
    rs = np.random.RandomState(seed)
    pool = rs.normal(0, 1, (D, n0 + 40 * frames)).astype(np.float32)
    pool /= np.linalg.norm(pool, axis=0)
    xy = rs.uniform(8, 300, (2, pool.shape[1]))
    seq = []
    for f in range(frames):
        if f == 3:
            seq.append((np.zeros((3, 0)), np.zeros((D, 0), np.float32)))
            continue
        if f == 5:
            seq.append((None, None))
            continue
        vis = np.sort(rs.permutation(n0 + 40 * f)[: n0 - 10 * f])
        d = pool[:, vis] + 0.06 * rs.normal(0, 1, (D, vis.size)).astype(np.float32)
        d = (d / np.linalg.norm(d, axis=0)).astype(np.float32)
        p = np.vstack((np.rint(xy[:, vis] + f * 1.5), rs.uniform(0.1, 1.0, (1, vis.size))))
        order = np.argsort(-p[2])
        seq.append((p[:, order].astype(np.float64), d[:, order]))
    return seq


def ros_stub_modules():
    """Import-time stand-ins for the ROS packages src/yolopoint_ros.py imports (rospy, rospkg, cv_bridge, message packages), so that
    the reference's own ``to_ros_msg`` can run here; the message classes are plain attribute bags."""
    import types

    class Bag:
        def __init__(self):
            self.instances = []

    mods = {}
    for name in ("rospy", "rospkg", "cv_bridge", "sensor_msgs", "sensor_msgs.msg", "object_instance_msgs", "object_instance_msgs.msg",
                 "keypoint_msg", "keypoint_msg.msg"):
        mods[name] = types.ModuleType(name)
    mods["sensor_msgs.msg"].Image = Bag
    mods["cv_bridge"].CvBridge = Bag
    mods["cv_bridge"].CvBridgeError = Exception
    mods["object_instance_msgs.msg"].ObjectInstance2D = type("ObjectInstance2D", (Bag,), {})
    mods["object_instance_msgs.msg"].ObjectInstance2DArray = type("ObjectInstance2DArray", (Bag,), {})
    mods["keypoint_msg.msg"].KeypointArray = type("KeypointArray", (Bag,), {})
    return mods


def reference_to_ros_msg(pts, desc, obj_preds, names):
    """Run the reference's YoloPointFrontendROS.to_ros_msg (src/yolopoint_ros.py:109-145) with stubbed ROS modules."""
    import types
    saved = {k: sys.modules.get(k) for k in ros_stub_modules()}
    sys.modules.update(ros_stub_modules())
    try:
        import yolopoint_ros
        fake = types.SimpleNamespace(names=names)
        return yolopoint_ros.YoloPointFrontendROS.to_ros_msg(fake, pts, desc, obj_preds, "hdr")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def golden_tracker(ns=None):
    """PointTracker.update / get_tracks (src/demo.py:358-441) over a synthetic sequence, and the KeypointArray / ObjectInstance2D
    flattening of to_ros_msg (src/yolopoint_ros.py:109-145), both produced by the reference's own code."""
    ns = ns or ref_import.load()
    import contextlib
    import io
    arrs = {}
    trk = ns.PointTracker(max_length=4, nn_thresh=0.7)
    for f, (p, d) in enumerate(tracker_sequence()):
        if p is not None:
            arrs[f"pts{f}"], arrs[f"desc{f}"] = p, d
            arrs[f"matches{f}"] = ns.PointTracker.nn_match_two_way(trk.last_desc if trk.last_desc is not None else np.zeros((d.shape[0], 0)), d, 0.7)
        with contextlib.redirect_stdout(io.StringIO()):
            trk.update(p, d)
        arrs[f"tracks{f}"] = trk.tracks.copy()
        arrs[f"count{f}"] = np.array(trk.track_count)
        arrs[f"long{f}"] = trk.get_tracks(2)
        arrs[f"offsets{f}"] = trk.get_offsets()
    arrs["none_frames"] = np.array([f for f, (p, _) in enumerate(tracker_sequence()) if p is None])
    # wire format
    rs = np.random.RandomState(4)
    pts = np.vstack((rs.randint(4, 636, 50), rs.randint(4, 476, 50), rs.uniform(0.1, 1, 50))).astype(np.float64)
    desc = rs.normal(0, 1, (64, 50)).astype(np.float32)
    det = torch.tensor([[10.7, 20.2, 110.9, 220.5, 0.91, 2.0], [300.1, 40.6, 420.4, 160.0, 0.55, 0.0], [5.5, 6.5, 7.5, 8.5, 0.41, 1.0]])
    names = ["car", "person", "bike"]
    km, am = reference_to_ros_msg(pts, desc, [det], names)
    arrs.update(w_pts=pts, w_desc=desc, w_det=det.numpy(), w_x=km.x, w_y=km.y, w_score=km.score, w_desc_len=np.asarray(km.desc_len),
                w_desc_flat=np.asarray(km.desc_flat), w_names=np.array(names),
                w_obj_index=np.array([m.class_index for m in am.instances]), w_obj_name=np.array([m.class_name for m in am.instances]),
                w_obj_prob=np.array([m.class_probabilities[0] for m in am.instances]),
                w_obj_box=np.array([[m.bounding_box_min_x, m.bounding_box_min_y, m.bounding_box_max_x, m.bounding_box_max_y] for m in am.instances]),
                w_obj_count=np.array([m.class_count for m in am.instances]))
    save("tracker.npz", **arrs)


def homography_batch(seed=0, B=6):
    """Normalised inverse homographies (rotation, scale, translation, perspective terms; the first one is the identity)."""
    rs = np.random.RandomState(seed)
    Hs = []
    for _ in range(B):
        a, sc = rs.uniform(-0.25, 0.25), rs.uniform(0.8, 1.2)
        Hs.append(np.array([[sc * np.cos(a), -sc * np.sin(a), rs.uniform(-.2, .2)], [sc * np.sin(a), sc * np.cos(a), rs.uniform(-.2, .2)],
                            [rs.uniform(-.15, .15), rs.uniform(-.15, .15), 1.0]], np.float32))
    Hs = np.stack(Hs)
    Hs[0] = np.eye(3)
    return Hs


def golden_homography(ns=None):
    """SURVEY.md section 8f rank 2 (contract pinned ahead of the kernel): warp_image_batch / compute_valid_mask (src/utils/utils.py:
    296-376) and the aggregation of src/export_homography.py:97-128, run by the reference's own functions."""
    ns = ns or ref_import.load()
    from utils.utils import compute_valid_mask, warp_image_batch
    B, H, W = 6, 48, 64
    rs = np.random.RandomState(5)
    heat = rs.rand(B, 1, H, W).astype(np.float32) ** 4
    Hs = homography_batch(0, B)
    th, tH = torch.from_numpy(heat), torch.from_numpy(Hs)
    mask = compute_valid_mask(torch.tensor([H, W]), tH)[:, None]
    wb = warp_image_batch(th, tH, mode="bilinear")
    wn = warp_image_batch(th, tH, mode="nearest")
    agg = torch.sum(warp_image_batch(th * mask, tH), dim=0) / torch.sum(warp_image_batch(mask, tH), dim=0)
    save("homography.npz", heat=heat, inv_homographies=Hs, mask=mask.numpy(), warp_bilinear=wb.numpy(), warp_nearest=wn.numpy(), aggregated=agg.numpy()[0])


def golden_infonce():
    """The descriptor loss of the reference's training script (src/train.py:8: infonce, src/utils/loss_functions.py:484-597) on the
    descriptor maps of losses.npz: 16 draws of its random pair / negative sampling."""
    ref_import.load()
    from utils.loss_functions import infonce
    g = np.load(os.path.join(OUT, "losses.npz"))
    d1, d2, mask, Hm = (torch.from_numpy(g[k]) for k in ("d1", "d2", "mask", "Hm"))
    np.random.seed(0)
    torch.manual_seed(0)
    vals = np.array([infonce(d1, d2, mask, Hm, num_samples_per_image=200, num_masked_non_matches_per_match=50).item() for _ in range(16)], np.float32)
    eye, ones = torch.eye(3).repeat(d1.shape[0], 1, 1), torch.ones_like(mask)
    same = np.array([infonce(d1, d1, ones, eye, num_samples_per_image=200, num_masked_non_matches_per_match=50).item() for _ in range(4)], np.float32)
    save("infonce.npz", linfonce=vals, linfonce_same=same)


def golden_losses():
    """Training losses (SURVEY.md section 8 row a11 consumers): the reference's own ComputeObjectLoss / ComputeDetectorLoss /
    descriptor_loss_sparse (src/utils/loss_functions.py:90-234, 600-619, 361-481) on seeded synthetic network outputs."""
    ns = ref_import.load()
    from utils.loss_functions import ComputeDetectorLoss, ComputeObjectLoss, descriptor_loss_sparse
    from utils.utils import getMasks, labels2Dto3D
    torch.manual_seed(0)
    m = ns.Model(names=NAMES80, version="n")
    cfg = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, iou_t=0.2, anchor_t=4.0, label_smoothing=0.0, fl_gamma=0.0)
    B, H, W = 2, 128, 160
    rs = np.random.RandomState(0)
    p = [torch.from_numpy(rs.randn(B, 3, H // s, W // s, 85).astype(np.float32)) for s in (8, 16, 32)]
    t = np.concatenate([np.stack([np.full(6, b), rs.randint(0, 80, 6), rs.uniform(.1, .9, 6), rs.uniform(.1, .9, 6), rs.uniform(.05, .5, 6),
                                  rs.uniform(.05, .5, 6)], 1) for b in range(B)]).astype(np.float32)
    for pi in p:
        pi.requires_grad_(True)
    lobj, items = ComputeObjectLoss(m, cfg, "cpu")(p, torch.from_numpy(t))
    lobj.backward()
    semi = torch.from_numpy(rs.randn(B, 65, H // 8, W // 8).astype(np.float32)).requires_grad_(True)
    lab = (rs.rand(B, 1, H, W) < 0.005).astype(np.float32)
    mask = np.ones((B, 1, H, W), np.float32)
    mask[:, :, :16] = 0
    ldet = ComputeDetectorLoss("cpu")(semi, labels2Dto3D(torch.from_numpy(lab)), getMasks(torch.from_numpy(mask), "cpu"))
    ldet.backward()
    d1 = torch.nn.functional.normalize(torch.from_numpy(rs.randn(B, 64, H // 8, W // 8).astype(np.float32)), dim=1)
    d2 = torch.nn.functional.normalize(d1 + 0.3 * torch.from_numpy(rs.randn(*d1.shape).astype(np.float32)), dim=1)
    Hm = torch.eye(3).repeat(B, 1, 1)
    Hm[1, 0, 2] = 0.1
    np.random.seed(0)
    ldesc = np.array([descriptor_loss_sparse(d1, d2, torch.from_numpy(mask), Hm, num_samples_per_image=200, num_masked_non_matches_per_match=50).item()
                      for _ in range(8)], np.float32)
    save("losses.npz", p0=p[0].detach().numpy(), p1=p[1].detach().numpy(), p2=p[2].detach().numpy(), targets=t, lobj=lobj.detach().numpy(),
         lobj_items=items.numpy(), g0=p[0].grad.numpy(), semi=semi.detach().numpy(), labels=lab, mask=mask, ldet=ldet.detach().numpy(),
         gsemi=semi.grad.numpy(), d1=d1.numpy(), d2=d2.numpy(), Hm=Hm.numpy(), ldesc=ldesc,
         labels3d=labels2Dto3D(torch.from_numpy(lab)).numpy(), mask3d=getMasks(torch.from_numpy(mask), "cpu").numpy())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "losses":
        os.makedirs(OUT, exist_ok=True)
        golden_losses()
    elif len(sys.argv) > 1 and sys.argv[1] == "v52":
        os.makedirs(OUT, exist_ok=True)
        golden_v52()
    elif len(sys.argv) > 1 and sys.argv[1] == "infonce":
        golden_infonce()
    elif len(sys.argv) > 1 and sys.argv[1] == "homography":
        os.makedirs(OUT, exist_ok=True)
        golden_homography()
    elif len(sys.argv) > 1 and sys.argv[1] == "tracker":
        os.makedirs(OUT, exist_ok=True)
        golden_tracker()
    else:
        main()
        golden_losses()
        golden_tracker()
        golden_homography()
        golden_infonce()
