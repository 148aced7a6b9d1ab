"""CPU oracle for the YOLOPoint hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a CPU restatement (numpy for integer/index work, torch-CPU fp32 for the
floating-point convolutions) of the reference algorithm of UniBwTAS/YOLOPoint for the one
hot path this repository accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package ``yolopoint_b200`` never does (it raises when its CUDA library is missing).

Parity pin: the reference repository ships no tests or golden vectors,
so every function here was pinned by running the *reference's own code* in the build
container (``oracle/make_golden.py``, which imports ``/root/reference/src``) and committing
the input/output pairs under ``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the
restatement against those vectors on every run, and ``tests/test_oracle_vs_reference.py``
re-checks it against the live reference whenever ``/root/reference`` is present.

Each function cites the reference file:line it restates (paths relative to the reference
repository root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# Network description (src/models/YOLOPoint.py:148-196, src/models/common.py:22-34,79-89,123-135,213-229)
# ----------------------------------------------------------------------------------------------

ANCHORS_DEFAULT = (  # src/models/YOLOPoint.py:11-15
    (10, 13, 16, 30, 33, 23),
    (30, 61, 62, 45, 59, 119),
    (116, 90, 156, 198, 373, 326),
)
VERSIONS = {  # src/models/YOLOPoint.py:36-45  -> (depth_multiple, width_multiple)
    "n": (0.33, 0.25), "s": (0.33, 0.5), "m": (0.67, 0.75), "l": (1.0, 1.0), "x": (1.33, 1.25),
}
BN_EPS = 1e-3  # src/models/common.py:18-20
STRIDES = (8.0, 16.0, 32.0)  # derived by the 256x256 dummy forward at src/models/YOLOPoint.py:64


def make_divisible(x: float, divisor: int) -> int:
    """src/utils/general_yolo.py:534"""
    return math.ceil(x / divisor) * divisor


def dims(version: str) -> Tuple[Tuple[int, ...], Tuple[int, ...]]:
    """Channel widths c1..c5 and depths n1..n3 (src/models/YOLOPoint.py:152-153)."""
    dm, wm = VERSIONS[version]
    cs = tuple(make_divisible(2 ** k * wm, 8) for k in range(6, 11))
    ns = tuple(max(round(k * dm), 1) for k in (3, 6, 9))
    return cs, ns


def _fold(sd: Dict[str, torch.Tensor], name: str, dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """Conv+BN folding exactly as src/utils/torch_utils_yolo.py:194-214 (fp32; ``dtype=torch.float64`` evaluates the same
    expressions in double precision, see OracleNet)."""
    w = sd[name + ".conv.weight"].to(dtype)
    if name + ".bn.weight" not in sd:  # already fused state dict
        return w, sd[name + ".conv.bias"].to(dtype)
    g, b = sd[name + ".bn.weight"].to(dtype), sd[name + ".bn.bias"].to(dtype)
    mu, var = sd[name + ".bn.running_mean"].to(dtype), sd[name + ".bn.running_var"].to(dtype)
    scale = g.div(torch.sqrt(BN_EPS + var))
    wf = torch.mm(torch.diag(scale), w.view(w.shape[0], -1)).view(w.shape)
    bf = b - g.mul(mu).div(torch.sqrt(var + BN_EPS))
    return wf, bf


class OracleNet:
    """Functional, BN-folded, eval-mode YOLOPoint forward (src/models/YOLOPoint.py:198-246).

    ``sd`` is a reference-format state dict (keys ``model.Conv1.conv.weight`` ...).  ``dtype=torch.float64`` evaluates the same
    graph on the same fp32 weights in double precision: the yardstick that tells how far ANY fp32 evaluation (the reference's
    own included) is from the exact result, used by the tests to set the float tolerances.
    """

    def __init__(self, sd: Dict[str, torch.Tensor], version: str, nc: int, model_name: str = "YOLOPoint", dtype=torch.float32):
        assert model_name in ("YOLOPoint", "YOLOPointv52"), model_name
        self.version, self.nc, self.no, self.model_name, self.dtype = version, nc, nc + 5, model_name, dtype
        sd = {(k[len("model."):] if k.startswith("model.") else k): v.detach().cpu() for k, v in sd.items()}
        self.sd = sd
        (self.c1, self.c2, self.c3, self.c4, self.c5), (self.n1, self.n2, self.n3) = dims(version)
        self._folded: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.anchors = sd["Detect.anchors"].to(dtype)  # (3,3,2), already divided by stride
        self.stride = torch.tensor(STRIDES, dtype=dtype)

    # -- building blocks -------------------------------------------------------------------
    def _conv(self, name: str, x: torch.Tensor, k: int, s: int, p: Optional[int] = None) -> torch.Tensor:
        """models/common.py:22-34 forward_fuse: act(conv(x)) with SiLU."""
        if name not in self._folded:
            self._folded[name] = _fold(self.sd, name, self.dtype)
        w, b = self._folded[name]
        p = k // 2 if p is None else p
        return F.silu(F.conv2d(x, w, b, stride=s, padding=p))

    def _bottleneck(self, name: str, x: torch.Tensor) -> torch.Tensor:
        """models/common.py:79-89 (shortcut=True, e=1.0 inside C3)."""
        return x + self._conv(name + ".cv2", self._conv(name + ".cv1", x, 1, 1), 3, 1)

    def _c3(self, name: str, x: torch.Tensor, n: int) -> torch.Tensor:
        """models/common.py:123-135."""
        y = self._conv(name + ".cv1", x, 1, 1)
        for i in range(n):
            y = self._bottleneck(f"{name}.m.{i}", y)
        return self._conv(name + ".cv3", torch.cat((y, self._conv(name + ".cv2", x, 1, 1)), 1), 1, 1)

    def _sppf(self, name: str, x: torch.Tensor) -> torch.Tensor:
        """models/common.py:213-229."""
        x = self._conv(name + ".cv1", x, 1, 1)
        y1 = F.max_pool2d(x, 5, 1, 2)
        y2 = F.max_pool2d(y1, 5, 1, 2)
        y3 = F.max_pool2d(y2, 5, 1, 2)
        return self._conv(name + ".cv2", torch.cat((x, y1, y2, y3), 1), 1, 1)

    def _c2f(self, name: str, x: torch.Tensor, n: int) -> torch.Tensor:
        """models/common.py:151-165 with Bottleneckv8 (:91-103): two 3x3 convs, e=1.0, shortcut=False (C2f's default)."""
        y = list(self._conv(name + ".cv1", x, 1, 1).chunk(2, 1))
        for i in range(n):
            y.append(self._conv(f"{name}.m.{i}.cv2", self._conv(f"{name}.m.{i}.cv1", y[-1], 3, 1), 3, 1))
        return self._conv(name + ".cv2", torch.cat(y, 1), 1, 1)

    def _detect(self, feats: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        """Detect.forward, eval branch (src/models/yolo.py:49-81)."""
        raw = []
        for i, t in enumerate(feats):
            t = F.conv2d(t, self.sd[f"Detect.m.{i}.weight"].to(self.dtype), self.sd[f"Detect.m.{i}.bias"].to(self.dtype))
            bs, _, ny, nx = t.shape
            raw.append(t.view(bs, 3, self.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous())
        return detect_decode(raw, self.anchors, self.stride), raw

    @torch.no_grad()
    def forward_v52(self, x: torch.Tensor, keep: Optional[dict] = None) -> Dict[str, object]:
        """YOLOPointv52.forward (src/models/YOLOPoint.py:295-342)."""
        up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
        x = self._conv("Conv1", x.to(self.dtype), 6, 2, 2)
        x = self._conv("Conv2", x, 3, 2)
        xa = self._c2f("Bottleneck1", x, self.n1)
        x = self._conv("Conv3", xa, 3, 2)
        semi = self._c2f("BottleneckDet", x, self.n1)
        xb = self._c2f("Bottleneck2", x, self.n2)
        descA = F.max_pool2d(xa, 2, 2)
        descB = up(self._conv("ConvDescB", xb, 3, 2, 1))
        desc = self._c2f("BottleneckDesc", torch.cat((descA, descB), 1), self.n1)
        dn = torch.norm(desc, p=2, dim=1)
        desc = desc.div(torch.unsqueeze(dn, 1))
        x = self._conv("Conv4", xb, 3, 2)
        xc = self._c2f("Bottleneck3", x, self.n3)
        x = self._conv("Conv5", xc, 3, 2)
        x = self._c2f("Bottleneck4", x, self.n1)
        xd = self._sppf("SPPooling", x)
        xe = self._c2f("Bottleneck5", torch.cat((up(xd), xc), 1), self.n1)
        xf = self._c2f("Bottleneck6", torch.cat((up(xe), xb), 1), self.n1)
        x = self._conv("Conv8", xf, 3, 2, 1)
        xg = self._c2f("Bottleneck7", torch.cat((x, xe), 1), self.n1)
        x = self._conv("Conv9", xg, 3, 2, 1)
        xh = self._c2f("Bottleneck8", torch.cat((x, xd), 1), self.n1)
        if keep is not None:
            keep.update(xa=xa, xb=xb, xc=xc, xd=xd, xe=xe, xf=xf, xg=xg, xh=xh)
        pred, raw = self._detect((xf, xg, xh))
        return {"semi": semi, "desc": desc, "objects": (pred, raw)}

    # -- network ---------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, keep: Optional[dict] = None) -> Dict[str, object]:
        if self.model_name == "YOLOPointv52":
            return self.forward_v52(x, keep)
        sd = self.sd
        up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
        x = self._conv("Conv1", x.to(self.dtype), 6, 2, 2)
        x = self._conv("Conv2", x, 3, 2)
        xa = self._c3("Bottleneck1", x, self.n1)
        x = self._conv("Conv3", xa, 3, 2)
        semi = self._c3("BottleneckDet", x, self.n1)
        semi = F.conv2d(semi, sd["ConvDet.weight"].to(self.dtype))
        xb = self._c3("Bottleneck2", x, self.n2)
        descA = self._conv("ConvDescA", xa, 3, 2, 1)
        descB = up(self._conv("ConvDescB", xb, 3, 2, 1))
        desc = self._c3("BottleneckDesc", torch.cat((descA, descB), 1), self.n1)
        desc = F.conv2d(desc, sd["ConvDesc.weight"].to(self.dtype), padding=1)
        dn = torch.norm(desc, p=2, dim=1)
        desc = desc.div(torch.unsqueeze(dn, 1))
        x = self._conv("Conv4", xb, 3, 2)
        xc = self._c3("Bottleneck3", x, self.n3)
        x = self._conv("Conv5", xc, 3, 2)
        x = self._c3("Bottleneck4", x, self.n1)
        x = self._sppf("SPPooling", x)
        xd = self._conv("Conv6", x, 1, 1, 0)
        x = self._c3("Bottleneck5", torch.cat((up(xd), xc), 1), self.n1)
        xe = self._conv("Conv7", x, 1, 1, 0)
        xf = self._c3("Bottleneck6", torch.cat((up(xe), xb), 1), self.n1)
        x = self._conv("Conv8", xf, 3, 2, 1)
        xg = self._c3("Bottleneck7", torch.cat((x, xe), 1), self.n1)
        x = self._conv("Conv9", xg, 3, 2, 1)
        xh = self._c3("Bottleneck8", torch.cat((x, xd), 1), self.n1)
        if keep is not None:
            keep.update(xa=xa, xb=xb, xc=xc, xd=xd, xe=xe, xf=xf, xg=xg, xh=xh)
        raw = []
        for i, t in enumerate((xf, xg, xh)):
            t = F.conv2d(t, sd[f"Detect.m.{i}.weight"].to(self.dtype), sd[f"Detect.m.{i}.bias"].to(self.dtype))
            bs, _, ny, nx = t.shape
            raw.append(t.view(bs, 3, self.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous())
        pred = detect_decode(raw, self.anchors, self.stride)
        return {"semi": semi, "desc": desc, "objects": (pred, raw)}


def detect_decode(raw: Sequence[torch.Tensor], anchors: torch.Tensor, stride: torch.Tensor) -> torch.Tensor:
    """Eval branch of Detect.forward / _make_grid (src/models/yolo.py:49-81).

    raw[i]: [B,3,ny,nx,no] logits.  Returns pred [B, A, no] with rows ordered level-major then
    (anchor, y, x); xy=(2*sig-0.5+grid)*stride, wh=(2*sig)^2*anchor*stride, rest=sig.
    """
    z = []
    for i, x in enumerate(raw):
        bs, na, ny, nx, no = x.shape
        yv, xv = torch.meshgrid(torch.arange(ny), torch.arange(nx), indexing="ij")
        grid = torch.stack((xv, yv), 2).expand(1, na, ny, nx, 2).to(x.dtype)
        ag = (anchors[i].clone() * stride[i]).view(1, na, 1, 1, 2).expand(1, na, ny, nx, 2).to(x.dtype)
        y = x.sigmoid()
        xy = (y[..., 0:2] * 2 - 0.5 + grid) * stride[i]
        wh = (y[..., 2:4] * 2) ** 2 * ag
        z.append(torch.cat((xy, wh, y[..., 4:]), -1).view(bs, -1, no))
    return torch.cat(z, 1)


# ----------------------------------------------------------------------------------------------
# Box NMS  (src/utils/general_yolo.py:124-235, 623-630; torchvision.ops.nms semantics)
# ----------------------------------------------------------------------------------------------

def nms_greedy(boxes: np.ndarray, scores: np.ndarray, iou_thres: float) -> np.ndarray:
    """Greedy IoU NMS with torchvision.ops.nms semantics (torchvision is an un-vendored,
    unpinned dependency of the reference, requirements.txt:20; container has 0.26.0).

    Published algorithm (torchvision/csrc/ops/cpu/nms_kernel.cpp): visit boxes by descending
    score (stable for ties), suppress j iff inter/(area_i+area_j-inter) > thr, all in fp32.
    Returns kept indices into ``boxes`` in visiting order.
    """
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    x1, y1, x2, y2 = (boxes[order, k] for k in range(4))
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, bool)
    keep = []
    thr = np.float32(iou_thres)
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(n):
            if suppressed[i]:
                continue
            keep.append(i)
            xx1 = np.maximum(x1[i], x1[i + 1:]); yy1 = np.maximum(y1[i], y1[i + 1:])
            xx2 = np.minimum(x2[i], x2[i + 1:]); yy2 = np.minimum(y2[i], y2[i + 1:])
            w = np.maximum(np.float32(0), xx2 - xx1); h = np.maximum(np.float32(0), yy2 - yy1)
            inter = w * h
            ovr = inter / (areas[i] + areas[i + 1:] - inter)
            suppressed[i + 1:] |= ovr > thr
    return order[np.asarray(keep, np.int64)]


def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    """src/utils/general_yolo.py:623-630 (fp32)."""
    y = np.empty_like(x)
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, labels=(), max_det=300) -> List[np.ndarray]:
    """src/utils/general_yolo.py:124-235 restated in numpy fp32 (nm=0, merge=False, no labels).

    Canonical tie order (the reference's ``argsort(descending=True)`` is unstable, SURVEY.md
    section 7): equal confidences keep candidate (row-major (box, class)) order.
    Returns a list (one per image) of float32 [n,6] arrays (x1,y1,x2,y2,conf,cls).
    """
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    pred = prediction.detach().cpu().numpy() if isinstance(prediction, torch.Tensor) else np.asarray(prediction)
    pred = pred.astype(np.float32, copy=False)
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}"
    assert not labels, "autolabelling path (general_yolo.py:171-178) is out of scope"
    bs, _, no = pred.shape
    nc = no - 5
    max_wh, max_nms = 7680, 30000
    multi_label = bool(multi_label) and nc > 1
    thr = np.float32(conf_thres)
    out = [np.zeros((0, 6), np.float32) for _ in range(bs)]
    for xi in range(bs):
        x = pred[xi]
        x = x[x[:, 4] > thr].copy()
        if not x.shape[0]:
            continue
        x[:, 5:] *= x[:, 4:5]
        box = xywh2xyxy(x[:, :4])
        if multi_label:
            i, j = np.nonzero(x[:, 5:] > thr)
            x = np.concatenate((box[i], x[i, 5 + j, None], j[:, None].astype(np.float32)), 1)
        else:
            j = x[:, 5:].argmax(1)
            conf = x[np.arange(x.shape[0]), 5 + j]
            x = np.concatenate((box, conf[:, None], j[:, None].astype(np.float32)), 1)[conf > thr]
        if classes is not None:
            x = x[(x[:, 5:6] == np.asarray(classes, np.float32)).any(1)]
        n = x.shape[0]
        if not n:
            continue
        x = x[np.argsort(-x[:, 4], kind="stable")[:max_nms]]
        c = x[:, 5:6] * np.float32(0 if agnostic else max_wh)
        keep = nms_greedy(x[:, :4] + c, x[:, 4], iou_thres)
        out[xi] = x[keep[:max_det]]
    return out


# ----------------------------------------------------------------------------------------------
# Keypoint heatmap + NMS  (src/utils/utils.py:118-182, 232-262, 465-485; src/demo.py:138-166)
# ----------------------------------------------------------------------------------------------

def flatten_detection(semi: np.ndarray, cell: int = 8, variant: str = "torch") -> np.ndarray:
    """[.., 65, Hc, Wc] logits -> [.., H, W] heatmap.

    variant "torch": ``flattenDetection`` (src/utils/utils.py:232-262) = softmax(dim=C), drop
    dustbin, PixelShuffle(8).  variant "demo": numpy form of src/demo.py:140-150 =
    exp(x) / (sum + 1e-5) with no max subtraction.  heat[8*hc+i, 8*wc+j] = p[8*i+j, hc, wc].
    """
    s = np.asarray(semi, np.float32)
    lead = s.shape[:-3]
    C, Hc, Wc = s.shape[-3:]
    assert C == cell * cell + 1
    s = s.reshape((-1, C, Hc, Wc))
    if variant == "torch":
        dense = torch.softmax(torch.from_numpy(s), dim=1).numpy()
    else:
        dense = np.exp(s)
        dense = dense / (np.sum(dense, axis=1, keepdims=True) + np.float32(.00001))
    nodust = dense[:, :-1].transpose(0, 2, 3, 1).reshape(-1, Hc, Wc, cell, cell)
    heat = nodust.transpose(0, 1, 3, 2, 4).reshape(-1, Hc * cell, Wc * cell)
    return heat.reshape(lead + (Hc * cell, Wc * cell))


def nms_fast(in_corners: np.ndarray, H: int, W: int, dist_thresh: int) -> Tuple[np.ndarray, np.ndarray]:
    """src/utils/utils.py:118-182, canonical (stable) tie order.  3xN (x, y, conf) in -> 3xK out."""
    grid = np.zeros((H, W), int)
    inds = np.zeros((H, W), int)
    inds1 = np.argsort(-in_corners[2, :], kind="stable")
    corners = in_corners[:, inds1]
    rcorners = corners[:2, :].round().astype(int)
    if rcorners.shape[1] == 0:
        return np.zeros((3, 0)).astype(int), np.zeros(0).astype(int)
    if rcorners.shape[1] == 1:
        return np.vstack((rcorners, in_corners[2])).reshape(3, 1), np.zeros((1)).astype(int)
    grid[rcorners[1], rcorners[0]] = 1
    inds[rcorners[1], rcorners[0]] = np.arange(rcorners.shape[1])
    pad = dist_thresh
    grid = np.pad(grid, ((pad, pad), (pad, pad)), mode="constant")
    for i in range(rcorners.shape[1]):
        px, py = rcorners[0, i] + pad, rcorners[1, i] + pad
        if grid[py, px] == 1:
            grid[py - pad:py + pad + 1, px - pad:px + pad + 1] = 0
            grid[py, px] = -1
    keepy, keepx = np.where(grid == -1)
    keepy, keepx = keepy - pad, keepx - pad
    inds_keep = inds[keepy, keepx]
    out = corners[:, inds_keep]
    inds2 = np.argsort(-out[-1, :], kind="stable")
    return out[:, inds2], inds1[inds_keep[inds2]]


def get_pts_from_heatmap(heatmap: np.ndarray, conf_thresh: float, nms_dist: int, border_remove: int = 4) -> np.ndarray:
    """src/utils/utils.py:465-485 == src/demo.py:151-166.  Returns float64 [3,N] (x, y, conf),
    confidence descending.  Canonical tie order: equal confidences in REVERSE raster order (what the
    reference's ascending argsort + [::-1] yields when its sorts behave stably)."""
    H, W = heatmap.shape
    ys, xs = np.where(heatmap >= conf_thresh)
    if len(ys) == 0:
        return np.zeros((3, 0))
    pts = np.zeros((3, len(ys)))
    pts[0, :] = xs
    pts[1, :] = ys
    pts[2, :] = heatmap[ys, xs]
    pts, _ = nms_fast(pts, H, W, dist_thresh=nms_dist)
    inds = np.argsort(pts[2, :], kind="stable")
    pts = pts[:, inds[::-1]].astype(np.float64)
    b = border_remove
    rm = (pts[0] < b) | (pts[0] >= W - b) | (pts[1] < b) | (pts[1] >= H - b)
    return pts[:, ~rm]


def get_pts_from_semi(semi: np.ndarray, conf_thresh=0.015, nms_dist=4) -> np.ndarray:
    """src/utils/utils.py:94-101."""
    return get_pts_from_heatmap(np.squeeze(flatten_detection(semi)), conf_thresh, nms_dist)


def filter_points_in_boxes(pts: np.ndarray, boxes: np.ndarray, H: int, W: int) -> np.ndarray:
    """Closure ``filter_points`` of src/demo.py:178-198 (no ROS template): drop keypoints that fall in
    any box, with np.rint + half-open python slices (negative indices wrap, as in the reference)."""
    mask = np.ones((H, W))
    points = pts.transpose()
    p2 = points[:, :2].astype(int)
    for row in np.asarray(boxes, np.float32):
        x0, y0, x1, y1 = np.rint(row[:4]).astype(int)
        mask[y0:y1, x0:x1] = 0
    return points[mask[p2[:, 1], p2[:, 0]] == 1].transpose()


# ----------------------------------------------------------------------------------------------
# Descriptor sampling + matching (src/demo.py:200-215, 300-341; evaluations/descriptor_evaluation.py:148-181)
# ----------------------------------------------------------------------------------------------

def sample_desc_from_points(coarse_desc: np.ndarray, pts: np.ndarray, cell: int = 8,
                            HW: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """Bilinear ``grid_sample(align_corners=True, zeros padding)`` of [1,D,Hc,Wc] descriptors at pixel
    coordinates, then column L2 normalisation; restates ATen's grid_sampler_2d CPU arithmetic in fp32.
    Returns float32 [D,N] (float64 zeros((D,0)) when N == 0, as src/demo.py:202-203)."""
    cd = np.asarray(coarse_desc, np.float32)
    cd = cd.reshape((1,) * (4 - cd.ndim) + cd.shape)[0]
    D, Hc, Wc = cd.shape
    H, W = HW if HW is not None else (Hc * cell, Wc * cell)
    if pts.shape[1] == 0:
        return np.zeros((D, 0))
    gx = (pts[0, :].astype(np.float64) / (float(W) / 2.) - 1.).astype(np.float32)
    gy = (pts[1, :].astype(np.float64) / (float(H) / 2.) - 1.).astype(np.float32)
    ix = ((gx + np.float32(1)) / np.float32(2)) * np.float32(Wc - 1)
    iy = ((gy + np.float32(1)) / np.float32(2)) * np.float32(Hc - 1)
    x0 = np.floor(ix); y0 = np.floor(iy)
    x1 = x0 + 1; y1 = y0 + 1
    wnw = (x1 - ix) * (y1 - iy); wne = (ix - x0) * (y1 - iy)
    wsw = (x1 - ix) * (iy - y0); wse = (ix - x0) * (iy - y0)
    out = np.zeros((D, pts.shape[1]), np.float32)
    for xs, ys, w in ((x0, y0, wnw), (x1, y0, wne), (x0, y1, wsw), (x1, y1, wse)):
        xi, yi = xs.astype(np.int64), ys.astype(np.int64)
        ok = (xi >= 0) & (xi < Wc) & (yi >= 0) & (yi < Hc)
        v = cd[:, np.clip(yi, 0, Hc - 1), np.clip(xi, 0, Wc - 1)]
        out += np.where(ok, v * w.astype(np.float32), np.float32(0))
    out /= np.linalg.norm(out, axis=0)[np.newaxis, :]
    return out


def nn_match_two_way(desc1: np.ndarray, desc2: np.ndarray, nn_thresh: float) -> np.ndarray:
    """src/demo.py:300-341 (same math as src/models/model_wrap.py:434-476).  [D,N1],[D,N2] -> [3,L]."""
    assert desc1.shape[0] == desc2.shape[0]
    if desc1.shape[1] == 0 or desc2.shape[1] == 0:
        return np.zeros((3, 0))
    if nn_thresh < 0.0:
        raise ValueError("'nn_thresh' should be non-negative")
    dmat = np.dot(desc1.T, desc2)
    dmat = np.sqrt(2 - 2 * np.clip(dmat, -1, 1))
    idx = np.argmin(dmat, axis=1)
    scores = dmat[np.arange(dmat.shape[0]), idx]
    keep = scores < nn_thresh
    idx2 = np.argmin(dmat, axis=0)
    keep = np.logical_and(keep, np.arange(len(idx)) == idx2[idx])
    matches = np.zeros((3, int(keep.sum())))
    matches[0, :] = np.arange(desc1.shape[1])[keep]
    matches[1, :] = idx[keep]
    matches[2, :] = scores[keep]
    return matches


# ----------------------------------------------------------------------------------------------
# Whole-frame pipeline (src/demo.py:125-230 without crop/resize bookkeeping)
# ----------------------------------------------------------------------------------------------

# ----------------------------------------------------------------------------------------------
# Homography adaptation (the next row; restated ahead of its kernel so that the contract is pinned)
# src/utils/utils.py:274-291 (warp_points), :333-376 (warp_image_batch); src/export_homography.py:92-128 (combine)
# ----------------------------------------------------------------------------------------------

def warp_image_batch(img: np.ndarray, mat_homo_inv: np.ndarray, mode: str = "bilinear") -> np.ndarray:
    """Inverse warp of a batch [B,C,H,W] by normalised homographies [B,3,3] (zeros padding, align_corners=True).

    Output pixel (y, x) samples the source at ``p = Hinv_b . (xn, yn, 1)``, ``(u, v) = p[:2] / p[2]`` in normalised coordinates,
    ``xn = linspace(-1, 1, W)[x]``, ``yn = linspace(-1, 1, H)[y]`` (torch.linspace values), i.e. source pixel
    ``ix = (u + 1) / 2 * (W - 1)``, ``iy = (v + 1) / 2 * (H - 1)``; bilinear weights and the four-term sum in ATen's order
    (nw, ne, sw, se), out-of-range corners contribute zero; ``mode='nearest'`` rounds half to even like ``nearbyint``."""
    img = np.asarray(img, np.float32)
    Hm = np.asarray(mat_homo_inv, np.float32).reshape(-1, 3, 3)
    B, Cc, H, W = img.shape
    xs = torch.linspace(-1, 1, W).numpy()
    ys = torch.linspace(-1, 1, H).numpy()
    xn, yn = np.meshgrid(xs, ys)                                   # [H,W] each
    pts = np.stack((xn.ravel(), yn.ravel(), np.ones(H * W, np.float32)), 0).astype(np.float32)   # [3, H*W]
    out = np.zeros_like(img)
    for b in range(B):
        p = (Hm[b] @ pts).astype(np.float32)
        u = (p[0] / p[2]).reshape(H, W)
        v = (p[1] / p[2]).reshape(H, W)
        ix = ((u + np.float32(1)) / np.float32(2)) * np.float32(W - 1)
        iy = ((v + np.float32(1)) / np.float32(2)) * np.float32(H - 1)
        if mode == "nearest":
            xi, yi = np.rint(ix).astype(np.int64), np.rint(iy).astype(np.int64)
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H) & np.isfinite(ix) & np.isfinite(iy)
            out[b] = np.where(ok, img[b][:, np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)], np.float32(0))
            continue
        x0, y0 = np.floor(ix), np.floor(iy)
        x1, y1 = x0 + 1, y0 + 1
        terms = ((x0, y0, (x1 - ix) * (y1 - iy)), (x1, y0, (ix - x0) * (y1 - iy)), (x0, y1, (x1 - ix) * (iy - y0)), (x1, y1, (ix - x0) * (iy - y0)))
        acc = np.zeros((Cc, H, W), np.float32)
        with np.errstate(invalid="ignore"):
            for xq, yq, w in terms:
                fin = np.isfinite(xq) & np.isfinite(yq)
                xi = np.where(fin, xq, -1).astype(np.int64)
                yi = np.where(fin, yq, -1).astype(np.int64)
                ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
                acc += np.where(ok, img[b][:, np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)] * w.astype(np.float32), np.float32(0))
        out[b] = acc
    return out


def homography_adaptation(heat: np.ndarray, valid_mask: np.ndarray, inv_homographies: np.ndarray) -> np.ndarray:
    """src/export_homography.py:97-128 without the letterbox slicing: heat [B,1,H,W] of the B warped copies of one image, valid masks
    [B,1,H,W], inverse homographies [B,3,3] -> the aggregated heatmap [H,W] = sum_b warp(heat_b * mask_b) / sum_b warp(mask_b)
    (0/0 = NaN where no copy covers a pixel, as in the reference)."""
    h = warp_image_batch(np.asarray(heat, np.float32) * np.asarray(valid_mask, np.float32), inv_homographies)
    m = warp_image_batch(valid_mask, inv_homographies)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (h.sum(0, dtype=np.float32) / m.sum(0, dtype=np.float32))[0]


DEFAULT_CFG = dict(  # configs/kitti_inference.yaml:5-16
    detection_threshold=0.12, nms=8, nn_thresh=0.7, conf_thres_box=0.4, iou_thres_box=0.45, max_det=1000,
)


def process_outputs(outs: Dict[str, object], H: int, W: int, cfg: dict = DEFAULT_CFG, filter_pts: bool = True,
                    heat_variant: str = "demo"):
    """Everything of process_img after the network: returns (pts[3,N] f64, desc[D,N] f32, boxes[n,6] f32)."""
    semi = outs["semi"].detach().cpu().numpy().squeeze(0)
    heat = flatten_detection(semi, variant=heat_variant)
    pts = get_pts_from_heatmap(heat, cfg["detection_threshold"], cfg["nms"])
    boxes = non_max_suppression(outs["objects"][0], cfg["conf_thres_box"], cfg["iou_thres_box"],
                                multi_label=True, agnostic=True, max_det=cfg["max_det"])[0]
    if pts.shape[1] == 0:
        return pts, None, boxes
    if filter_pts:
        pts = filter_points_in_boxes(pts, boxes, H, W)
    desc = sample_desc_from_points(outs["desc"].detach().cpu().numpy(), pts, HW=(H, W))
    return pts, desc, boxes


def process_frame(net: OracleNet, frame_u8: np.ndarray, cfg: dict = DEFAULT_CFG, filter_pts: bool = True):
    """uint8 [H,W,3] frame (H,W multiples of 32) -> (pts, desc, boxes); src/demo.py:126-215."""
    H, W, _ = frame_u8.shape
    inp = torch.from_numpy(frame_u8.transpose((2, 0, 1)).astype(np.float32) / 255.).unsqueeze(0)
    outs = net.forward(inp)
    return process_outputs(outs, H, W, cfg, filter_pts)
