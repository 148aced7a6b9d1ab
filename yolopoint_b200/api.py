"""Reference-named entry points of the hot path (same names, argument meaning, return types and error
behaviour as the reference's Python functions), executed by the B200 kernels.

  non_max_suppression      src/utils/general_yolo.py:124-235
  xywh2xyxy                src/utils/general_yolo.py:623-630 (host helper, tiny)
  flattenDetection         src/utils/utils.py:232-262
  getPtsFromHeatmap        src/utils/utils.py:465-485
  getPtsFromSemi           src/utils/utils.py:94-101
  nms_fast                 src/utils/utils.py:118-182
  sample_desc_from_points  src/evaluations/descriptor_evaluation.py:148-181
  nn_match_two_way         src/demo.py:300-341 (PointTracker.nn_match_two_way)
  warp_image_batch         src/utils/utils.py:333-376
  homography_adaptation    src/export_homography.py:97-128 (heat-map aggregation over the warped copies of one image)
  detect / extract_keypoints / match   convenience names requested by BASELINE.json north_star

numpy in -> numpy out where the reference does so; every call moves data to the current CUDA device, runs
the kernels and (where the reference returns host data) reads the result back.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("yolopoint_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(x, dtype=torch.float32) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.detach().to(device=x.device if x.is_cuda else _dev(), dtype=dtype)
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(_dev())


def xywh2xyxy(x):
    y = x.clone() if isinstance(x, torch.Tensor) else np.copy(x)
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300, nm=0, cap: Optional[int] = None) -> List[torch.Tensor]:
    """Returns a list (one per image) of [n,6] tensors (x1,y1,x2,y2,conf,cls) on the prediction's device."""
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if nm != 0 or (labels and any(len(l) for l in labels)):
        raise NotImplementedError("mask outputs (nm>0) and autolabelling (labels) are outside the accelerated hot path")
    src_dev = prediction.device if isinstance(prediction, torch.Tensor) else torch.device("cpu")
    pred = _to_dev(prediction)
    B, A, no = pred.shape
    cap = cap or 30016   # >= the reference's max_nms = 30000: the kernel then handles any candidate count like the reference does
    while True:
        boxes, count = ops.box_nms(pred, conf_thres, iou_thres, multi_label, agnostic, max_det, classes, cap=cap)
        cnt = count.cpu().numpy()
        if (cnt >= 0).all():
            break
        # only reachable with an explicit cap < max_nms: grow to what the kernel reported and redo (never truncate silently)
        need = int((-1 - cnt[cnt < 0]).max())
        cap = min(30016, max(2 * cap, (need + 63) // 64 * 64))
    out = []
    for b in range(B):
        n = int(cnt[b])
        out.append(boxes[b, :n].clone().to(src_dev) if src_dev.type == "cuda" else boxes[b, :n].cpu())
    return out


def detect(pred_or_model_out, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, max_det=300):
    """decode+NMS convenience: accepts Model.forward()'s dict, its 'objects' tuple, or the pred tensor."""
    p = pred_or_model_out
    if isinstance(p, dict):
        p = p["objects"]
    return non_max_suppression(p, conf_thres, iou_thres, classes=classes, agnostic=agnostic, multi_label=multi_label, max_det=max_det)


def flattenDetection(semi, cell_size=8):
    """[65,Hc,Wc] -> [1,H,W];  [B,65,Hc,Wc] -> [B,1,H,W]  (torch tensor on the input's device, as the reference)."""
    assert cell_size == 8, "the kernel is specialised for the reference's fixed cell size 8 (src/demo.py:27)"
    src_dev = semi.device if isinstance(semi, torch.Tensor) else torch.device("cpu")
    s = _to_dev(semi)
    batch = s.dim() == 4
    if not batch:
        s = s.unsqueeze(0)
    heat = ops.heatmap(s.contiguous(), "nchw", variant=0)
    heat = heat.unsqueeze(1) if batch else heat
    return heat if src_dev.type == "cuda" else heat.cpu()


def _pts_to_numpy(pts: torch.Tensor, count: torch.Tensor, b: int = 0, what: str = "max_pts") -> np.ndarray:
    n = int(count[b].item())
    if n < 0:
        raise RuntimeError(f"keypoint buffer overflow: {-1 - n} points; raise {what}")
    return pts[b, :n].double().cpu().numpy().T.copy() if n else np.zeros((3, 0))


def getPtsFromHeatmap(heatmap, conf_thresh, nms_dist, max_pts: Optional[int] = None) -> np.ndarray:
    """HxW heatmap -> float64 [3,N] (x, y, conf), confidence descending, 4-px border removed."""
    h = _to_dev(heatmap)
    assert h.dim() == 2
    H, W = h.shape
    if max_pts is None:  # survivors are at least nms_dist+1 apart in Chebyshev distance
        max_pts = ((H + nms_dist) // (nms_dist + 1)) * ((W + nms_dist) // (nms_dist + 1))
    pts, count = ops.keypoints(h.unsqueeze(0), conf_thresh, nms_dist, border=4, max_pts=max(int(max_pts), 1))
    return _pts_to_numpy(pts, count)


def getPtsFromSemi(semi, conf_thresh=0.015, nms_dist=4) -> np.ndarray:
    heat = flattenDetection(_to_dev(semi))
    return getPtsFromHeatmap(heat.reshape(heat.shape[-2], heat.shape[-1]), conf_thresh, nms_dist)


def nms_fast(in_corners, H, W, dist_thresh):
    """Greedy grid NMS of 3xN (x, y, conf) corners -> (3xK survivors by confidence desc, K input indices).

    Corners are rounded to pixels as in the reference; if several corners round to the same pixel the
    reference keeps the attributes of the lowest-confidence duplicate while ranking by the highest --
    that quirk is reproduced by resolving duplicates on the host before the kernel runs."""
    c = np.asarray(in_corners, dtype=np.float64)
    if c.shape[1] == 0:
        return np.zeros((3, 0)).astype(int), np.zeros(0).astype(int)
    r = c[:2].round().astype(int)
    if c.shape[1] == 1:
        return np.vstack((r, c[2])).reshape(3, 1), np.zeros((1)).astype(int)
    order = np.argsort(-c[2], kind="stable")
    lin = r[1, order] * W + r[0, order]
    first = {}   # pixel -> confidence that ranks it (first visit = highest confidence)
    last = {}    # pixel -> input index whose attributes are reported (last write to `inds`)
    for pos, (pix, idx) in enumerate(zip(lin.tolist(), order.tolist())):
        first.setdefault(pix, c[2, idx])
        last[pix] = idx
    heat = torch.full((H * W,), float("-inf"), dtype=torch.float32)
    pix = torch.tensor(list(first.keys()), dtype=torch.long)
    heat[pix] = torch.tensor(list(first.values()), dtype=torch.float32)
    max_pts = ((H + dist_thresh) // (dist_thresh + 1)) * ((W + dist_thresh) // (dist_thresh + 1))
    pts, count = ops.keypoints(heat.view(1, H, W).to(_dev()), -3.0e38, dist_thresh, border=0, max_pts=max(max_pts, 1))
    p = _pts_to_numpy(pts, count)
    idx = np.array([last[int(y) * W + int(x)] for x, y in zip(p[0], p[1])], dtype=int)
    out = c[:, idx].copy()
    out[:2] = r[:, idx]
    return out, idx


def sample_desc_from_points(coarse_desc, pts, device=None, cell_size=8) -> np.ndarray:
    """coarse_desc [D,Hc,Wc] or [1,D,Hc,Wc]; pts [3,N] or [2,N] (x, y[, conf]) -> float32 [D,N] unit descriptors."""
    cd = _to_dev(coarse_desc)
    if cd.dim() != 4:
        cd = cd.view(*(1,) * (4 - cd.dim()), *cd.shape)
    D, Hc, Wc = cd.shape[1:]
    pts = np.asarray(pts)
    if pts.ndim != 2 or pts.shape[1] == 0:
        return np.empty((D, 0))
    N = pts.shape[1]
    p = torch.zeros((1, N, 3), dtype=torch.float32)
    p[0, :, :2] = torch.from_numpy(np.ascontiguousarray(pts[:2].T.astype(np.float32)))
    out = ops.sample_desc(cd.contiguous(), p.to(cd.device), None, (Hc * cell_size, Wc * cell_size), "nchw")
    return out[0].T.contiguous().cpu().numpy()


def nn_match_two_way(desc1, desc2, nn_thresh) -> np.ndarray:
    """desc1 [D,N1], desc2 [D,N2] unit descriptors -> float64 [3,L] rows (i, j, score), ascending i."""
    d1s = desc1.shape
    d2s = desc2.shape
    assert d1s[0] == d2s[0]
    if d1s[1] == 0 or d2s[1] == 0:
        return np.zeros((3, 0))
    if nn_thresh < 0.0:
        raise ValueError("'nn_thresh' should be non-negative")
    D = d1s[0]
    Dp = (D + 3) // 4 * 4
    a = _to_dev(desc1).T.contiguous()
    b = _to_dev(desc2).T.contiguous()
    if Dp != D:
        a = torch.nn.functional.pad(a, (0, Dp - D))
        b = torch.nn.functional.pad(b, (0, Dp - D))
    m, cnt = ops.match_two_way(a, None, b, None, float(nn_thresh))
    n = int(cnt.item())
    return m[:n].double().cpu().numpy().T.copy() if n else np.zeros((3, 0))


match = nn_match_two_way


def extract_keypoints(semi, desc, conf_thresh=0.015, nms_dist=4, boxes=None):
    """semi [65,Hc,Wc] / [1,65,Hc,Wc], desc [D,Hc,Wc] / [1,D,Hc,Wc] -> (pts float64 [3,N], desc float32 [D,N]):
    heatmap -> threshold -> NMS -> border filter [-> in-box filter] -> descriptor sampling, all on the device."""
    s = _to_dev(semi)
    s = s.unsqueeze(0) if s.dim() == 3 else s
    d = _to_dev(desc)
    d = d.unsqueeze(0) if d.dim() == 3 else d
    assert s.shape[0] == 1, "extract_keypoints works per frame (as getPtsFromSemi does)"
    heat = ops.heatmap(s.contiguous(), "nchw", 0)
    H, W = heat.shape[1:]
    max_pts = ((H + nms_dist) // (nms_dist + 1)) * ((W + nms_dist) // (nms_dist + 1))
    bx = bc = None
    if boxes is not None and len(boxes):
        bx = _to_dev(boxes).reshape(1, -1, 6).contiguous()
        bc = torch.tensor([bx.shape[1]], dtype=torch.int32, device=bx.device)
    pts, count = ops.keypoints(heat, conf_thresh, nms_dist, 4, bx, bc, max_pts=max_pts)
    out = ops.sample_desc(d.contiguous(), pts, count, (H, W), "nchw")
    n = int(count.item())
    if n < 0:
        raise RuntimeError("keypoint buffer overflow")
    return (pts[0, :n].double().cpu().numpy().T.copy() if n else np.zeros((3, 0))), out[0, :n].T.contiguous().cpu().numpy()


def warp_image_batch(img, mat_homo_inv, device="cpu", mode="bilinear", padding_mode="zeros"):
    """Inverse warp of a batch of images (src/utils/utils.py:333-376): ``img`` [B,1,H,W] / [B,C,H,W], [H,W,3] (RGB), [B,H,W,3] or
    [H,W]; ``mat_homo_inv`` [B,3,3] or [3,3].  ``device`` is accepted for signature compatibility; the warp always runs on the
    current CUDA device and the result comes back where the input was."""
    if padding_mode != "zeros":
        raise NotImplementedError("the reference only ever warps with padding_mode='zeros'")
    src_dev = img.device if isinstance(img, torch.Tensor) else torch.device("cpu")
    t = _to_dev(img)
    transposed = False
    if t.shape[-1] == 3 and t.dim() == 3:       # the reference unsqueezes and hands [1,H,W,3] to grid_sample as [N,C,H,W] = [1,H,W,3]
        t = t.unsqueeze(0)
    elif t.shape[-1] == 3 and t.dim() == 4:
        transposed = True
        t = t.transpose(1, 3).transpose(2, 3)
    elif t.dim() in (2, 3):
        t = t.reshape(1, 1, t.shape[-2], t.shape[-1]) if t.dim() == 2 else t.reshape(1, 1, t.shape[0], t.shape[1])
    hm = _to_dev(mat_homo_inv).reshape(-1, 3, 3)
    out = ops.warp_batch(t.contiguous(), hm, mode)
    if transposed:
        out = out.transpose(1, 3).transpose(1, 2)
    return out if src_dev.type == "cuda" else out.cpu()


def homography_adaptation(heatmap, mask_2D, inv_homographies, pad=None):
    """The aggregation step of the homography-adaptation export (src/export_homography.py:97-128): ``heatmap`` [B,1,H,W] =
    flattenDetection(semi) of the B warped copies of one image, ``mask_2D`` [B,1,H,W] their valid masks, ``inv_homographies``
    [B,3,3] -> ``sum_b warp(heatmap_b * mask_b) / sum_b warp(mask_b)`` as [1,H',W'] (``outputs`` of the reference; NaN where no copy
    covers a pixel), fused into one kernel.  ``pad`` = the sample's letterbox tuple: the reference's slicing (:103-109, including its
    use of the WIDTH for the row bound) is applied to the result, which is equivalent because the aggregation is per pixel."""
    src_dev = heatmap.device if isinstance(heatmap, torch.Tensor) else torch.device("cpu")
    h, m = _to_dev(heatmap), _to_dev(mask_2D)
    B, H, W = h.shape[0], h.shape[-2], h.shape[-1]
    out = ops.homography_adapt(h.reshape(B, H, W), m.reshape(B, H, W), _to_dev(inv_homographies).reshape(-1, 3, 3)).unsqueeze(0)
    if pad:
        height, width = H, W
        if pad[1]:
            out = out[:, int(pad[0]):width - int(pad[1]), :]
        if pad[3]:
            out = out[:, :, int(pad[2]):height - int(pad[3])]
    return out if src_dev.type == "cuda" else out.cpu()
