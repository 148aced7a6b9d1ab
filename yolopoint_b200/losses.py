"""Training losses of the reference, restated for the training-step path.

These are the consumers of ``Model.forward`` in train mode (src/train.py:208-241): they are small-tensor PyTorch code in the
reference as well (row f3 lists fusing them into kernels as later work), so they stay PyTorch here and run on whatever
device the network outputs live on.  Same names / arguments / return values as the reference, so ``train.py`` can import
them from here:

  ComputeObjectLoss     src/utils/loss_functions.py:90-234   (YOLOv5 box / objectness / class loss, CIoU)
  bbox_iou              src/utils/metrics_yolo.py:202-240
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

import torch
import torch.nn as nn
import torch.nn.functional as F


def bbox_iou(box1, box2, xywh=True, CIoU=False, eps=1e-7):
    """IoU / complete-IoU of row-aligned boxes [n,4] (metrics_yolo.py:202-240)."""
    if xywh:
        (x1, y1, w1, h1), (x2, y2, w2, h2) = box1.chunk(4, 1), box2.chunk(4, 1)
        b1x1, b1x2, b1y1, b1y2 = x1 - w1 / 2, x1 + w1 / 2, y1 - h1 / 2, y1 + h1 / 2
        b2x1, b2x2, b2y1, b2y2 = x2 - w2 / 2, x2 + w2 / 2, y2 - h2 / 2, y2 + h2 / 2
    else:
        b1x1, b1y1, b1x2, b1y2 = box1.chunk(4, 1)
        b2x1, b2y1, b2x2, b2y2 = box2.chunk(4, 1)
        w1, h1, w2, h2 = b1x2 - b1x1, b1y2 - b1y1, b2x2 - b2x1, b2y2 - b2y1
    inter = (torch.min(b1x2, b2x2) - torch.max(b1x1, b2x1)).clamp(0) * (torch.min(b1y2, b2y2) - torch.max(b1y1, b2y1)).clamp(0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    if not CIoU:
        return iou
    cw = torch.max(b1x2, b2x2) - torch.min(b1x1, b2x1)
    ch = torch.max(b1y2, b2y2) - torch.min(b1y1, b2y1)
    c2 = cw ** 2 + ch ** 2 + eps
    rho2 = ((b2x1 + b2x2 - b1x1 - b1x2) ** 2 + (b2y1 + b2y2 - b1y1 - b1y2) ** 2) / 4
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / (h2 + eps)) - torch.atan(w1 / (h1 + eps)), 2)
    with torch.no_grad():
        alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


def smooth_BCE(eps=0.1):
    return 1.0 - 0.5 * eps, 0.5 * eps


class ComputeObjectLoss:
    """loss_functions.py:90-234 (focal loss / autobalance variants are configuration the shipped YAMLs leave off)."""

    def __init__(self, model, config, device, autobalance=False):
        if config.get("fl_gamma", 0.0) > 0:
            raise NotImplementedError("fl_gamma > 0 (focal loss) is not used by the reference configs")
        self.BCEcls = nn.BCEWithLogitsLoss(pos_weight=torch.tensor([config["cls_pw"]], device=device))
        self.BCEobj = nn.BCEWithLogitsLoss(pos_weight=torch.tensor([config["obj_pw"]], device=device))
        self.cp, self.cn = smooth_BCE(eps=config.get("label_smoothing", 0.0))
        m = getattr(model, "module", model).model.Detect
        self.balance = {3: [4.0, 1.0, 0.4]}.get(m.nl, [4.0, 1.0, 0.25, 0.06, 0.02])
        self.gr, self.hyp = 1.0, config
        self.na, self.nc, self.nl, self.anchors, self.device = m.na, m.nc, m.nl, m.anchors, device

    def __call__(self, p, targets, built=None):
        """``built`` = a precomputed ``build_targets(p, targets)`` (it only needs the shapes of ``p``, so a training step can run it
        -- and its host synchronisations -- before the forward pass is launched)."""
        dev = self.device
        lcls, lbox, lobj = (torch.zeros(1, device=dev) for _ in range(3))
        tcls, tbox, indices, anchors = built if built is not None else self.build_targets(p, targets)
        for i, pi in enumerate(p):
            b, a, gj, gi = indices[i]
            tobj = torch.zeros(pi.shape[:4], dtype=pi.dtype, device=dev)
            n = b.shape[0]
            if n:
                pxy, pwh, _, pcls = pi[b, a, gj, gi].split((2, 2, 1, self.nc), 1)
                pxy = pxy.sigmoid() * 2 - 0.5
                pwh = (pwh.sigmoid() * 2) ** 2 * anchors[i]
                iou = bbox_iou(torch.cat((pxy, pwh), 1), tbox[i], CIoU=True).squeeze()
                lbox = lbox + (1.0 - iou).mean()
                iou = iou.detach().clamp(0).type(tobj.dtype)
                if self.gr < 1:
                    iou = (1.0 - self.gr) + self.gr * iou
                tobj[b, a, gj, gi] = iou
                if self.nc > 1:
                    t = torch.full_like(pcls, self.cn, device=dev)
                    t[range(n), tcls[i]] = self.cp
                    lcls = lcls + self.BCEcls(pcls, t)
            lobj = lobj + self.BCEobj(pi[..., 4], tobj) * self.balance[i]
        lbox = lbox * self.hyp["box"]
        lobj = lobj * self.hyp["obj"]
        lcls = lcls * self.hyp["cls"]
        return (lbox + lobj + lcls), torch.cat((lbox, lobj, lcls)).detach()

    def build_targets(self, p, targets):
        """targets [n,6] = (image, class, x, y, w, h) normalised -> per level (class, box, indices, anchors): every target is
        assigned to the anchors within ratio anchor_t and to the cell it falls in plus its two nearest neighbours."""
        dev = self.device
        na, nt = self.na, targets.shape[0]
        tcls, tbox, indices, anch = [], [], [], []
        gain = torch.ones(7, device=dev)
        ai = torch.arange(na, device=dev).float().view(na, 1).repeat(1, nt)
        targets = torch.cat((targets.repeat(na, 1, 1), ai[..., None]), 2)
        g = 0.5
        off = torch.tensor([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]], device=dev).float() * g
        for i in range(self.nl):
            anchors, shape = self.anchors[i], p[i].shape
            gain[2:6] = torch.tensor(shape)[[3, 2, 3, 2]]
            t = targets * gain
            if nt:
                r = t[..., 4:6] / anchors[:, None]
                j = torch.max(r, 1 / r).max(2)[0] < self.hyp["anchor_t"]
                t = t[j]
                gxy = t[:, 2:4]
                gxi = gain[[2, 3]] - gxy
                j, k = ((gxy % 1 < g) & (gxy > 1)).T
                l, m = ((gxi % 1 < g) & (gxi > 1)).T
                j = torch.stack((torch.ones_like(j), j, k, l, m))
                t = t.repeat((5, 1, 1))[j]
                offsets = (torch.zeros_like(gxy)[None] + off[:, None])[j]
            else:
                t = targets[0]
                offsets = 0
            bc, gxy, gwh, a = t.chunk(4, 1)
            a, (b, c) = a.long().view(-1), bc.long().T
            gij = (gxy - offsets).long()
            gi, gj = gij.T
            indices.append((b, a, gj.clamp_(0, shape[2] - 1), gi.clamp_(0, shape[3] - 1)))
            tbox.append(torch.cat((gxy - gij, gwh), 1))
            anch.append(anchors[a])
            tcls.append(c)
        return tcls, tbox, indices, anch


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


class ComputeDetectorLoss:
    """loss_functions.py:600-619: BCE between softmax(semi) and the 65-channel labels, summed over channels, masked mean."""

    def __init__(self, device):
        self.device = device

    def __call__(self, inp, target, mask):
        loss = F.binary_cross_entropy(torch.softmax(inp, dim=1), target, reduction="none")
        loss = (loss.sum(dim=1) * mask).sum()
        return loss / (mask.sum() + 1e-10)


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
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            neg = (da @ db.t()).gather(1, rnd.t()).t()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
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
