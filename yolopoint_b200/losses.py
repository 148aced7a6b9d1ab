"""Training losses of the reference, restated for the training-step path.

These are the consumers of ``Model.forward`` in train mode (src/train.py:208-241): they are small-tensor PyTorch code in the
reference as well (row f3 lists fusing them into kernels as later work), so they stay PyTorch here and run on whatever
device the network outputs live on.  Same names / arguments / return values as the reference, so ``train.py`` can import
them from here:

  ComputeObjectLoss     src/utils/loss_functions.py:90-234   (YOLOv5 box / objectness / class loss, CIoU) -- re-designed without
                        boolean indexing: fixed-shape masked candidates, no host synchronisation (see the class docstring)
  bbox_iou, ciou_xywh   src/utils/metrics_yolo.py:202-240
  ComputeDetectorLoss   src/utils/loss_functions.py:600-619  (65-way cell classification, BCE on the softmax)
  labels2Dto3D/getMasks src/utils/utils.py:184-209, 103-116
  descriptor_loss_sparse src/utils/loss_functions.py:361-481 (sampled hinge loss between frame / warped-frame descriptors)

One deliberate difference: the negative-sample indices of ``descriptor_loss_sparse`` are drawn with ``torch.randint`` on the
tensor's device instead of ``numpy.random.randint`` on the host (the reference forces a host round trip per step there,
loss_functions.py:451-466); the distribution is the same, the random stream is not.  On CUDA the sampled negative similarities
are gathered from one ``da @ db^T`` GEMM instead of a materialised ``[K, n, D]`` product (same values up to TF32 rounding).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def ciou_xywh(p: torch.Tensor, t: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """Complete-IoU of row-aligned centre-format boxes p, t [..., 4] = (x, y, w, h) -> [...].

    CIoU = IoU - rho^2 / c^2 - alpha v (Zheng et al. 2020, the form src/utils/metrics_yolo.py:202-240 evaluates with CIoU=True):
    rho = centre distance, c = diagonal of the smallest enclosing box, v = (4 / pi^2)(atan(w_t / h_t) - atan(w_p / h_p))^2,
    alpha = v / (v - IoU + 1 + eps) treated as a constant in the backward pass.  ``eps`` enters where the reference adds it
    (union, c^2, the two aspect ratios, alpha), so the values agree to rounding."""
    px, py, pw, ph = p.unbind(-1)
    tx, ty, tw, th = t.unbind(-1)
    p_l, p_r, p_t, p_b = px - pw / 2, px + pw / 2, py - ph / 2, py + ph / 2
    t_l, t_r, t_t, t_b = tx - tw / 2, tx + tw / 2, ty - th / 2, ty + th / 2
    inter = (torch.minimum(p_r, t_r) - torch.maximum(p_l, t_l)).clamp(0) * (torch.minimum(p_b, t_b) - torch.maximum(p_t, t_t)).clamp(0)
    iou = inter / (pw * ph + tw * th - inter + eps)
    diag2 = (torch.maximum(p_r, t_r) - torch.minimum(p_l, t_l)) ** 2 + (torch.maximum(p_b, t_b) - torch.minimum(p_t, t_t)) ** 2 + eps
    rho2 = ((t_l + t_r - p_l - p_r) ** 2 + (t_t + t_b - p_t - p_b) ** 2) / 4
    v = (4 / math.pi ** 2) * (torch.atan(tw / (th + eps)) - torch.atan(pw / (ph + eps))) ** 2
    alpha = (v / (v - iou + (1 + eps))).detach()
    return iou - (rho2 / diag2 + v * alpha)


def bbox_iou(box1, box2, xywh=True, CIoU=False, eps=1e-7):
    """Reference-named entry point (src/utils/metrics_yolo.py:202-240) for the two forms the hot path uses: row-aligned [n,4]
    boxes -> [n,1] IoU or CIoU."""
    if not xywh:
        (l1, t1, r1, b1), (l2, t2, r2, b2) = box1.unbind(-1), box2.unbind(-1)
        box1 = torch.stack(((l1 + r1) / 2, (t1 + b1) / 2, r1 - l1, b1 - t1), -1)
        box2 = torch.stack(((l2 + r2) / 2, (t2 + b2) / 2, r2 - l2, b2 - t2), -1)
    if CIoU:
        return ciou_xywh(box1, box2, eps).unsqueeze(-1)
    px, py, pw, ph = box1.unbind(-1)
    tx, ty, tw, th = box2.unbind(-1)
    inter = (torch.minimum(px + pw / 2, tx + tw / 2) - torch.maximum(px - pw / 2, tx - tw / 2)).clamp(0) * \
            (torch.minimum(py + ph / 2, ty + th / 2) - torch.maximum(py - ph / 2, ty - th / 2)).clamp(0)
    return (inter / (pw * ph + tw * th - inter + eps)).unsqueeze(-1)


class TargetPlan:
    """Everything ``ComputeObjectLoss`` needs from the labels, per Detect level, as FIXED-SHAPE tensors over all
    E = 5 x anchors x targets assignment candidates with a validity mask (no boolean indexing, so building and using it never
    waits for the device): flat cell index of the candidate, the target box relative to its cell, its anchor, its class."""

    def __init__(self, levels):
        self.levels = levels      # list of dicts: valid [E] bool, cell [E] long (b, a, gj, gi flattened), tbox [E,4], anchor [E,2], cls [E] long


class ComputeObjectLoss:
    """YOLOv5's box / objectness / class loss as the reference configures it (src/utils/loss_functions.py:90-234: CIoU box loss, BCE
    objectness against the detached CIoU, BCE classes; label smoothing, focal loss and autobalance are options the shipped YAMLs
    leave off).

    Target assignment (the rule of :120-234): a target is a candidate for an anchor of a level when no side of the target is more
    than ``anchor_t`` times longer or shorter than the anchor's; it is assigned to the cell that contains its centre and to the
    (up to two) neighbouring cells its centre is closest to.  The reference expresses this with boolean indexing, which on a GPU
    makes the host wait for the device five times per level; here all 5 x na x nt candidates are kept with a validity mask and the
    sums are masked, which gives the same loss values (tests/test_losses.py, against vectors of the reference) without any
    synchronisation, so the whole loss can be enqueued behind the forward pass.  Where several targets claim one (image, anchor,
    cell), the reference's ``tobj[b, a, gj, gi] = iou`` keeps the last one in its candidate order (offset variant, anchor,
    target); the same order decides here."""

    def __init__(self, model, config, device, autobalance=False):
        if config.get("fl_gamma", 0.0) > 0 or autobalance:
            raise NotImplementedError("focal loss / autobalance are not used by the reference configs")
        self.hyp, self.device = config, device
        self.cls_pw = torch.tensor([config["cls_pw"]], device=device)
        self.obj_pw = torch.tensor([config["obj_pw"]], device=device)
        smooth = config.get("label_smoothing", 0.0)
        self.cp, self.cn = 1.0 - 0.5 * smooth, 0.5 * smooth        # positive / negative class targets
        det = getattr(model, "module", model).model.Detect
        self.na, self.nc, self.nl, self.anchors = det.na, det.nc, det.nl, det.anchors
        self.balance = [4.0, 1.0, 0.4] if det.nl == 3 else [4.0, 1.0, 0.25, 0.06, 0.02]
        self.gr = 1.0
        self.fused = True           # CUDA predictions: csrc/object_loss.cu (no fallback: a missing library raises)
        self.anchors_host = [[float(v) for v in a.flatten().tolist()] for a in det.anchors.detach().cpu()]   # per level, in cells (read once)

    # -- labels only ---------------------------------------------------------------------------
    def build_targets(self, p, targets) -> TargetPlan:
        """``p``: the per-level predictions [B, na, ny, nx, no] (only their shapes are used; meta tensors are fine);
        ``targets`` [nt, 6] = (image, class, x, y, w, h), box normalised to [0, 1]."""
        dev, na = self.device, self.na
        targets = targets.to(dev).float()
        nt = targets.shape[0]
        img, cls = targets[:, 0].long(), targets[:, 1].long()
        half = torch.tensor([[0.0, 0.0], [0.5, 0.0], [0.0, 0.5], [-0.5, 0.0], [0.0, -0.5]], device=dev)     # centre, left, up, right, down
        levels = []
        for i in range(self.nl):
            B, _, ny, nx, _ = p[i].shape
            size = torch.tensor([nx, ny], device=dev, dtype=torch.float32)
            anchors = self.anchors[i].to(dev).float()                                   # [na, 2] in cells
            gxy, gwh = targets[:, 2:4] * size, targets[:, 4:6] * size                  # [nt, 2] in cells
            ratio = gwh[None] / anchors[:, None]                                       # [na, nt, 2]
            shape_ok = torch.maximum(ratio, 1 / ratio).amax(2) < self.hyp["anchor_t"]  # [na, nt]
            low = (gxy % 1 < 0.5) & (gxy > 1)                                          # centre in the left / upper half of its cell
            inv = size - gxy
            high = (inv % 1 < 0.5) & (inv > 1)                                         # ... in the right / lower half
            near = torch.stack((torch.ones(nt, dtype=torch.bool, device=dev), low[:, 0], low[:, 1], high[:, 0], high[:, 1]))   # [5, nt]
            valid = near[:, None, :] & shape_ok[None]                                  # [5, na, nt]
            cell = (gxy[None] - half[:, None]).long()                                  # [5, nt, 2] (x, y), truncation like the reference
            gi, gj = cell[..., 0].clamp(0, nx - 1), cell[..., 1].clamp(0, ny - 1)
            a_idx = torch.arange(na, device=dev)
            flat = ((img[None, None] * na + a_idx[None, :, None]) * ny + gj[:, None]) * nx + gi[:, None]      # [5, na, nt]
            tbox = torch.cat((gxy[None] - cell.float(), gwh[None].expand(5, nt, 2)), 2)  # [5, nt, 4] relative to the UNclamped cell
            E = 5 * na * nt
            levels.append(dict(valid=valid.reshape(E), cell=flat.reshape(E), tbox=tbox[:, None].expand(5, na, nt, 4).reshape(E, 4),
                               anchor=anchors[None, :, None].expand(5, na, nt, 2).reshape(E, 2), cls=cls[None, None].expand(5, na, nt).reshape(E),
                               cells=B * na * ny * nx))
        return TargetPlan(levels)

    # -- predictions -----------------------------------------------------------------------------
    def __call__(self, p, targets, built: Optional[TargetPlan] = None):
        """-> (loss [1], detached (box, obj, cls) terms [3]).  ``built`` = a precomputed ``build_targets(p, targets)``.

        On a CUDA device the whole loss over all levels and its gradient are four kernel launches (csrc/object_loss.cu,
        ``fused=True``, the default there); ``fused=False`` / CPU tensors run the PyTorch statement of the same arithmetic below,
        which is what the CPU tests pin to the reference's vectors and what the GPU tests check the kernels against."""
        if self.fused and all(pi.is_cuda for pi in p):
            # a precomputed plan is used as given; otherwise the target assignment runs inside the kernels from the label list
            out4 = _ObjectLossFused.apply(self, built, None if built is not None else targets, *p)
            return out4[0:1], out4[1:4].detach()
        plan = built if built is not None else self.build_targets(p, targets)
        return self._call_torch(p, plan)

    def _call_torch(self, p, plan: TargetPlan):
        dev = self.device
        lbox = torch.zeros(1, device=dev)
        lobj = torch.zeros(1, device=dev)
        lcls = torch.zeros(1, device=dev)
        for i, pi in enumerate(p):
            lv = plan.levels[i]
            valid, w = lv["valid"], lv["valid"].to(pi.dtype)
            rows = pi.reshape(-1, pi.shape[-1])
            obj_logit = rows[:, 4]
            tobj = torch.zeros(lv["cells"] + 1, dtype=pi.dtype, device=dev)          # last slot swallows the invalid candidates
            if valid.numel():
                n = w.sum()
                q = rows[lv["cell"]]                                                  # [E, no] (invalid candidates read a real row; masked below)
                xy = q[:, 0:2].sigmoid() * 2 - 0.5
                wh = (q[:, 2:4].sigmoid() * 2) ** 2 * lv["anchor"]
                ciou = ciou_xywh(torch.cat((xy, wh), 1), lv["tbox"])
                lbox = lbox + ((1.0 - ciou) * w).sum() / n.clamp(min=1)
                score = ciou.detach().clamp(0).to(tobj.dtype)
                if self.gr < 1:
                    score = (1.0 - self.gr) + self.gr * score
                # the last valid candidate of a cell (in candidate order) sets its objectness target
                order = torch.arange(valid.numel(), device=dev)
                slot = torch.where(valid, lv["cell"], torch.full_like(lv["cell"], lv["cells"]))
                owner = torch.full((lv["cells"] + 1,), -1, dtype=torch.long, device=dev).scatter_reduce(0, slot, order, "amax")
                mine = valid & (owner[slot] == order)
                tobj = tobj.scatter(0, torch.where(mine, slot, torch.full_like(slot, lv["cells"])), torch.where(mine, score, torch.zeros_like(score)))
                if self.nc > 1:
                    tcls = torch.full_like(q[:, 5:], self.cn)
                    tcls.scatter_(1, lv["cls"][:, None], self.cp)
                    bce = F.binary_cross_entropy_with_logits(q[:, 5:], tcls, pos_weight=self.cls_pw, reduction="none")
                    lcls = lcls + (bce * w[:, None]).sum() / (n * self.nc).clamp(min=1)
            lobj = lobj + F.binary_cross_entropy_with_logits(obj_logit, tobj[:-1], pos_weight=self.obj_pw) * self.balance[i]
        lbox, lobj, lcls = lbox * self.hyp["box"], lobj * self.hyp["obj"], lcls * self.hyp["cls"]
        return lbox + lobj + lcls, torch.cat((lbox, lobj, lcls)).detach()


class _ObjectLossFused(torch.autograd.Function):
    """ComputeObjectLoss over all Detect levels: loss terms and d loss / d predictions from ``yp_object_loss`` (claim / candidate /
    cells / finalize kernels, csrc/object_loss.cu), with the target assignment evaluated inside the kernels from the label list
    ``targets`` [nt,6] (``plan`` None) or read from a precomputed ``TargetPlan``.  -> [4] = (loss, box, obj, cls)."""

    @staticmethod
    def forward(ctx, crit, plan, targets, *p):
        import ctypes as C
        from . import _lib
        L = _lib.lib(require_device=True)
        dev = p[0].device
        nl, no = len(p), p[0].shape[-1]
        assert nl <= 5 and all(pi.dtype == torch.float32 and pi.shape[-1] == no for pi in p), [(pi.dtype, pi.shape) for pi in p]
        keep = []                                                   # device buffers the launches read: alive until enqueued
        levels = (_lib.YpObjLossLevel * nl)()
        grads = []
        if plan is None:
            tg = targets.to(device=dev, dtype=torch.float32).contiguous()
            assert tg.dim() == 2 and tg.shape[1] == 6, tuple(tg.shape)
            keep.append(tg)
        for i, pi in enumerate(p):
            rows = pi.detach().contiguous()
            d = torch.empty_like(rows)
            keep.append(rows)
            grads.append(d)
            levels[i].pred, levels[i].dpred = rows.data_ptr(), d.data_ptr()
            levels[i].cells, levels[i].balance = rows.numel() // no, float(crit.balance[i])
            if plan is None:
                B, na, ny, nx, _ = pi.shape
                assert na == crit.na and 2 * na <= 16
                levels[i].targets = tg.data_ptr() if tg.shape[0] else None
                levels[i].nt, levels[i].na, levels[i].nx, levels[i].ny, levels[i].nb = tg.shape[0], na, nx, ny, B
                levels[i].E = 5 * na * tg.shape[0]
                for k, v in enumerate(crit.anchors_host[i]):
                    levels[i].anchors[k] = v
                continue
            lv = plan.levels[i]
            E = int(lv["valid"].numel())
            valid = lv["valid"].to(device=dev, dtype=torch.bool).contiguous()
            cell = lv["cell"].to(device=dev, dtype=torch.int64).contiguous()
            tbox = lv["tbox"].to(device=dev, dtype=torch.float32).contiguous()
            anchor = lv["anchor"].to(device=dev, dtype=torch.float32).contiguous()
            cls = lv["cls"].to(device=dev, dtype=torch.int64).contiguous()
            keep += [valid, cell, tbox, anchor, cls]
            levels[i].valid, levels[i].cell, levels[i].tbox = valid.data_ptr(), cell.data_ptr(), tbox.data_ptr()
            levels[i].anchor, levels[i].cls, levels[i].E = anchor.data_ptr(), cls.data_ptr(), E
            assert int(lv["cells"]) == rows.numel() // no
        hp = _lib.YpObjLossParams(cp=crit.cp, cn=crit.cn, cls_pw=float(crit.hyp["cls_pw"]), obj_pw=float(crit.hyp["obj_pw"]), gr=crit.gr,
                                  w_box=float(crit.hyp["box"]), w_obj=float(crit.hyp["obj"]), w_cls=float(crit.hyp["cls"]), eps=1e-7,
                                  anchor_t=float(crit.hyp["anchor_t"]))
        ws = torch.empty(max(int(L.yp_object_loss_workspace_bytes(levels, nl)), 16), dtype=torch.uint8, device=dev)
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        _lib.check(L.yp_object_loss(levels, nl, no, crit.nc, C.byref(hp), out4.data_ptr(), ws.data_ptr(), ws.numel(),
                                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        ctx.save_for_backward(*grads)
        return out4

    @staticmethod
    def backward(ctx, grad_out):
        g = grad_out[0]                                             # only the total (out4[0]) carries gradient; the terms are reports
        return (None, None, None) + tuple(d * g for d in ctx.saved_tensors)


def labels2Dto3D(labels, cell_size=8, add_dustbin=True):
    """[B,1,H,W] keypoint map -> [B,64(+1),Hc,Wc] cell-major labels, normalised over channels (utils.py:184-209)."""
    B, _, H, W = labels.shape
    labels = F.pixel_unshuffle(labels, cell_size)
    if add_dustbin:
        dustbin = 1 - labels.sum(dim=1)
        dustbin[dustbin < 1.0] = 0
        labels = torch.cat((labels, dustbin.view(B, 1, H // cell_size, W // cell_size)), dim=1)
        labels = labels.div(labels.sum(dim=1).unsqueeze(1))
    return labels


def getMasks(mask_2D, device, cell_size=8):
    """[B,1,H,W] valid mask -> [B,Hc,Wc]: a cell is valid iff all of its pixels are (utils.py:103-116)."""
    return torch.prod(labels2Dto3D(mask_2D.to(device), cell_size=cell_size, add_dustbin=False).float(), 1)


class _DetectorLossFused(torch.autograd.Function):
    """labels2Dto3D + getMasks + ComputeDetectorLoss, forward and backward in one pass over the logits (csrc/loss.cu)."""

    @staticmethod
    def forward(ctx, semi, labels_2D, mask_2D):
        import ctypes as C
        from . import _lib
        L = _lib.lib(require_device=True)
        B, Cc, Hc, Wc = semi.shape
        assert Cc == 65 and semi.dtype == torch.float32, (semi.shape, semi.dtype)
        lab = labels_2D.to(semi.device).float().contiguous().view(B, Hc * 8, Wc * 8)
        msk = mask_2D.to(semi.device).float().contiguous().view(B, Hc * 8, Wc * 8)
        dsemi = torch.empty_like(semi)                       # same memory format as the logits (channels-last in the training step)
        assert dsemi.stride() == semi.stride()
        out2 = torch.empty(2, dtype=torch.float32, device=semi.device)
        ws = torch.empty(L.yp_detector_loss_workspace_bytes(B, Hc, Wc), dtype=torch.uint8, device=semi.device)
        sB, sC, sH, sW = semi.stride()
        _lib.check(L.yp_detector_loss(semi.data_ptr(), sB, sC, sH, sW, lab.data_ptr(), msk.data_ptr(), B, Hc, Wc, dsemi.data_ptr(), out2.data_ptr(),
                                      ws.data_ptr(), ws.numel(), C.c_void_p(torch.cuda.current_stream(semi.device).cuda_stream)))
        ctx.save_for_backward(dsemi)
        return out2[0]

    @staticmethod
    def backward(ctx, grad_out):
        (dsemi,) = ctx.saved_tensors
        return dsemi * grad_out, None, None


class ComputeDetectorLoss:
    """loss_functions.py:600-619: BCE between softmax(semi) and the 65-channel labels, summed over channels, masked mean.

    ``__call__(inp, target, mask)`` is the reference's call on prepared [B,65,Hc,Wc] labels / [B,Hc,Wc] cell masks (PyTorch ops, any
    device).  ``from_2d(inp, labels_2D, mask_2D)`` takes the [B,1,H,W] keypoint map and valid mask the data loader delivers and, on a
    CUDA device, runs label preparation, loss and its gradient as one fused kernel (same value, tests/test_gpu_train.py)."""

    def __init__(self, device):
        self.device = device

    def __call__(self, inp, target, mask):
        loss = F.binary_cross_entropy(torch.softmax(inp, dim=1), target, reduction="none")
        loss = (loss.sum(dim=1) * mask).sum()
        return loss / (mask.sum() + 1e-10)

    def from_2d(self, inp, labels_2D, mask_2D):
        if inp.is_cuda:
            return _DetectorLossFused.apply(inp.float(), labels_2D, mask_2D)
        return self(inp, labels2Dto3D(labels_2D.to(inp.device)), getMasks(mask_2D, inp.device))


def _warp_points(points, homographies):
    """points [N,2] (x,y), homographies [B,3,3] -> [B,N,2]  (utils.py:274-290)."""
    ones = torch.ones((points.shape[0], 1), device=points.device, dtype=points.dtype)
    ph = torch.cat((points, ones), 1)
    w = homographies @ ph.t().unsqueeze(0)
    w = w.transpose(2, 1)
    return w[:, :, :2] / w[:, :, 2:]


def _warp_mask_nearest(mask, inv_h):
    """Inverse-warp a [B,1,H,W] mask with homographies given in normalised [-1,1] coordinates (utils.py:333-376)."""
    B, _, H, W = mask.shape
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, H, device=mask.device), torch.linspace(-1, 1, W, device=mask.device), indexing="ij")
    grid = _warp_points(torch.stack((xs, ys), 2).view(-1, 2), inv_h).view(B, H, W, 2).float()
    return F.grid_sample(mask, grid, mode="nearest", align_corners=True, padding_mode="zeros")


@torch.no_grad()
def descriptor_pairs(mask_valid_warp, inv_homographies, B, Hc, Wc, num_samples_per_image=1500, num_masked_non_matches_per_match=120, cell_size=8,
                     device="cpu"):
    """The sampling half of descriptor_loss_sparse (loss_functions.py:374-427, 451-466): positive pairs (grid_sample coordinates in the
    frame and in the warped frame) and the indices of the random negatives.  It depends on the masks / homographies only -- not on
    the network outputs -- so a training step can run it (it synchronises with the host to size the sample pool) before the forward
    passes are launched."""
    ys, xs = torch.meshgrid(torch.arange(Hc, device=device), torch.arange(Wc, device=device), indexing="ij")
    uv_a = torch.stack((xs.reshape(-1), ys.reshape(-1)), 1).float()
    inv_h = inv_homographies.to(device).float()
    valid = getMasks(_warp_mask_nearest(mask_valid_warp.to(device).float(), inv_h), device, cell_size)
    valid = (valid == 1.0).flatten(1, -1)
    trans = torch.tensor([[2.0 / Wc, 0.0, -1.0], [0.0, 2.0 / Hc, -1.0], [0.0, 0.0, 1.0]], dtype=torch.float32, device=device)
    uv_b = _warp_points(uv_a, trans.inverse() @ inv_h @ trans).round_()
    pool = min(num_samples_per_image, int(valid.sum(1).min()))
    pa, pb = [], []
    for b in range(B):
        idx = valid[b].nonzero().squeeze(1)
        idx = idx[torch.randperm(idx.shape[0], device=device)[:pool]]
        pa.append(uv_a[idx])
        pb.append(uv_b[b][idx])
    scale = torch.tensor([Wc, Hc], dtype=torch.float32, device=device)
    pa = torch.stack(pa) / scale * 2 - 1
    pb = torch.stack(pb) / scale * 2 - 1
    n, K = B * pool, num_masked_non_matches_per_match
    rnd = torch.randint(0, n, (K, n), device=device)
    same = rnd == torch.arange(n, device=device).unsqueeze(0)
    rnd = torch.where(same, (rnd + 1 + torch.randint(0, max(n - 1, 1), (K, n), device=device)) % n, rnd)   # never the match itself
    return pa, pb, rnd


class _SimilarityTF32(torch.autograd.Function):
    """a @ b^T with TF32 tensor cores for the forward GEMM AND the two backward GEMMs (24000 x 24000 x 256 each in the training
    step: 10.7 ms on the fp32 SIMT path, 1 ms on tensor cores), with ``torch.backends.cuda.matmul.allow_tf32`` switched on only
    around these three matmuls -- the process-wide setting is left as the caller had it."""

    @staticmethod
    def _mm(x, y):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            return x @ y
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _SimilarityTF32._mm(a, b.t())

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return _SimilarityTF32._mm(g, b), _SimilarityTF32._mm(g.t(), a)


def _pair_similarities(descriptors, descriptors_warped, pairs):
    """Cosine similarities of the sampled pairs: pos [n] (a_i . b_i) and neg [K, n] (a_i . b_{rnd[k, i]}), n = B * pool
    (loss_functions.py:429-471 / 553-591: the part the two descriptor losses share)."""
    pa, pb, rnd = pairs

    def sample(desc, pts):
        return F.grid_sample(desc, pts.unsqueeze(1), mode="bilinear", align_corners=True).squeeze(2).transpose(1, 2)

    da = sample(descriptors, pa)
    db = sample(descriptors_warped, pb)
    pos = (da * db).sum(-1).flatten()
    da, db = da.flatten(0, 1), db.flatten(0, 1)
    n = da.shape[0]
    if da.is_cuda and n <= 40000:
        # The reference materialises db[rnd] as a [K, n, D] tensor (4.9 GB for 8 x 3000 samples, K = 200, D = 256) and multiplies it
        # by the broadcast queries (loss_functions.py:468-471, 587-591).  The same K x n similarities are entries of the n x n matrix
        # da @ db^T: one GEMM (TF32 tensor cores, 2.3 GB result) + a gather, ~5x less memory traffic forward and backward.
        neg = _SimilarityTF32.apply(da, db).gather(1, rnd.t()).t()
    else:
        neg = (da.unsqueeze(0) * db[rnd]).sum(-1)
    return pos, neg


def descriptor_loss_sparse(descriptors, descriptors_warped, mask_valid_warp, inv_homographies, num_samples_per_image=1500,
                           num_masked_non_matches_per_match=120, cell_size=8, device="cpu", pairs=None):
    """loss_functions.py:361-481 (the hinge form; the training script itself uses ``infonce`` below).  Positive pairs: every valid cell
    centre of the frame and its (rounded) position in the warped frame, a random subset of equal size per image; loss = mean hinge
    (1 - <a,b>) over the pairs + mean over the violating ones of the hinge (<a, b'> - 0.1) against random other warped samples.
    ``pairs`` = a precomputed ``descriptor_pairs(...)`` result."""
    device = descriptors.device
    B, _, Hc, Wc = descriptors.shape
    assert Hc * Wc >= num_samples_per_image, "Number of samples per image must be greater than number of pixels in image"
    if pairs is None:
        pairs = descriptor_pairs(mask_valid_warp, inv_homographies, B, Hc, Wc, num_samples_per_image, num_masked_non_matches_per_match, cell_size, device)
    pos, neg = _pair_similarities(descriptors, descriptors_warped, pairs)
    neg = torch.clamp(neg - 0.1, min=0).flatten()
    neg_loss = neg.sum() / (torch.count_nonzero(neg) + 1)
    return torch.clamp(1 - pos, min=0).mean() + neg_loss


def infonce(descriptors, descriptors_warped, mask_valid_warp, inv_homographies, num_samples_per_image=1500, num_masked_non_matches_per_match=120,
            cell_size=8, device="cpu", tau=0.07, pairs=None):
    """loss_functions.py:484-597 -- the descriptor loss the reference's training script uses (src/train.py:8: ``from
    utils.loss_functions import infonce as descriptor_loss_sparse``).  Same sampled pairs as the hinge form; every match is one
    (1 + K)-way classification: logits = [<a_i, b_i>, <a_i, b_rnd(1,i)>, ..., <a_i, b_rnd(K,i)>] / tau, loss = mean cross entropy
    of the match against its K random non-matches."""
    device = descriptors.device
    B, _, Hc, Wc = descriptors.shape
    assert Hc * Wc >= num_samples_per_image, "Number of samples per image must be greater than number of pixels in image"
    if pairs is None:
        pairs = descriptor_pairs(mask_valid_warp, inv_homographies, B, Hc, Wc, num_samples_per_image, num_masked_non_matches_per_match, cell_size, device)
    pos, neg = _pair_similarities(descriptors, descriptors_warped, pairs)
    logits = torch.cat((pos.unsqueeze(1), neg.t()), dim=1) / tau
    return -F.log_softmax(logits, dim=1)[:, 0].mean()
