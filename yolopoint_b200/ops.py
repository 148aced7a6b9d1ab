"""Device-side post-processing operators (thin wrappers over the C ABI; torch tensors carry the memory).

All functions take and return CUDA tensors, enqueue on the current torch stream and never synchronise;
variable-length results come back as (max-size buffer, int32 count) pairs.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import YpNmsParams

_ws_cache = {}


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _workspace(dev, kind: str, nbytes: int) -> torch.Tensor:
    """Scratch buffer per (device, kind, stream): calls on different streams never share scratch, and a buffer that is replaced
    by a larger one is kept alive by the caching allocator until the work already enqueued on its stream has finished."""
    stream = torch.cuda.current_stream(dev)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), kind, stream.cuda_stream)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        if t is not None:
            t.record_stream(stream)
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        _ws_cache[key] = t
    return t


_const_cache: dict = {}


def _const_i32(dev, value: int) -> torch.Tensor:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), int(value))
    t = _const_cache.get(key)
    if t is None:
        t = torch.tensor([int(value)], dtype=torch.int32, device=dev)
        _const_cache[key] = t
    return t


def _need_cuda(t: torch.Tensor, what: str):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"{what} must be a CUDA tensor: yolopoint_b200 has no CPU path")


def class_mask_tensor(classes: Optional[Sequence[int]], nc: int, dev) -> Optional[torch.Tensor]:
    if classes is None:
        return None
    words = [0] * ((nc + 31) // 32)
    for c in classes:
        c = int(c)
        if 0 <= c < nc:
            words[c >> 5] |= 1 << (c & 31)
    return torch.tensor([w - (1 << 32) if w >= (1 << 31) else w for w in words], dtype=torch.int32, device=dev)


def box_nms(pred: torch.Tensor, conf_thres: float, iou_thres: float, multi_label: bool, agnostic: bool, max_det: int,
            classes: Optional[Sequence[int]] = None, cap: int = 30016, max_nms: int = 30000, max_wh: float = 7680.0,
            out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, class_mask: Optional[torch.Tensor] = None):
    """pred [B,A,no] fp32 -> (boxes [B,max_det,6], count int32 [B]).  With cap >= max_nms any number of candidates is handled like
    the reference (the max_nms most confident enter the NMS); count < 0 (= -1-n) only with an explicit cap < max_nms that overflowed."""
    _need_cuda(pred, "prediction")
    L = _lib.lib(require_device=True)
    pred = pred.contiguous().float()
    B, A, no = pred.shape
    dev = pred.device
    cap = (int(cap) + 63) // 64 * 64
    if out is None:
        out = (torch.zeros((B, max_det, 6), dtype=torch.float32, device=dev), torch.zeros((B,), dtype=torch.int32, device=dev))
    boxes, count = out
    nbytes = L.yp_box_nms_workspace_bytes(B, A, no, cap)
    ws = _workspace(dev, "nms", nbytes)
    if class_mask is None:
        class_mask = class_mask_tensor(classes, no - 5, dev)
    p = YpNmsParams(float(conf_thres), float(iou_thres), int(bool(multi_label)), int(bool(agnostic)), int(max_det), int(max_nms),
                    float(max_wh), class_mask.data_ptr() if class_mask is not None else None)
    _lib.check(L.yp_box_nms(pred.data_ptr(), B, A, no, C.byref(p), cap, boxes.data_ptr(), count.data_ptr(), ws.data_ptr(), ws.numel(),
                            _stream(dev)))
    return boxes, count


def heatmap(semi: torch.Tensor, layout: str = "nchw", variant: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """semi: [B,65,Hc,Wc] (layout 'nchw') or [B,Hc,Wc,C>=65] ('nhwc') fp32 logits -> heat [B,8Hc,8Wc]."""
    _need_cuda(semi, "semi")
    L = _lib.lib(require_device=True)
    assert semi.dtype == torch.float32 and semi.dim() == 4
    if layout == "nchw":
        B, Cc, Hc, Wc = semi.shape
        sB, sC, sH, sW = semi.stride()
    else:
        B, Hc, Wc, Cc = semi.shape
        sB, sH, sW, sC = semi.stride()
    assert Cc >= 65
    if out is None:
        out = torch.empty((B, Hc * 8, Wc * 8), dtype=torch.float32, device=semi.device)
    _lib.check(L.yp_heatmap(semi.data_ptr(), B, Hc, Wc, sB, sC, sH, sW, int(variant), out.data_ptr(), _stream(semi.device)))
    return out


def keypoints(heat: torch.Tensor, conf_thresh: float, nms_dist: int, border: int = 4, boxes: Optional[torch.Tensor] = None,
              box_count: Optional[torch.Tensor] = None, max_pts: int = 8192, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """heat [B,H,W] fp32 -> (pts [B,max_pts,3] (x,y,conf) confidence-descending, count int32 [B])."""
    _need_cuda(heat, "heatmap")
    L = _lib.lib(require_device=True)
    heat = heat.contiguous()
    B, H, W = heat.shape
    dev = heat.device
    if out is None:
        out = (torch.zeros((B, max_pts, 3), dtype=torch.float32, device=dev), torch.zeros((B,), dtype=torch.int32, device=dev))
    pts, count = out
    nbytes = L.yp_keypoints_workspace_bytes(B, H, W, max_pts)
    ws = _workspace(dev, "kp", nbytes)
    bptr = cptr = None
    box_ld = 0
    if boxes is not None:
        assert boxes.is_contiguous() and boxes.dim() == 3 and boxes.shape[2] == 6 and box_count is not None
        bptr, cptr, box_ld = boxes.data_ptr(), box_count.data_ptr(), boxes.shape[1]
    _lib.check(L.yp_keypoints(heat.data_ptr(), B, H, W, float(conf_thresh), int(nms_dist), int(border), bptr, cptr, box_ld,
                              pts.data_ptr(), count.data_ptr(), max_pts, ws.data_ptr(), ws.numel(), _stream(dev)))
    return pts, count


def sample_desc(desc: torch.Tensor, pts: torch.Tensor, count: Optional[torch.Tensor], img_hw: Tuple[int, int], layout: str = "nchw",
                D: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """desc [B,D,Hc,Wc] ('nchw') or [B,Hc,Wc,D] ('nhwc'); pts [B,N,3] -> [B,N,D] unit descriptors (rows >= count untouched)."""
    _need_cuda(desc, "coarse_desc")
    L = _lib.lib(require_device=True)
    assert desc.dtype == torch.float32 and pts.dtype == torch.float32 and pts.is_contiguous()
    if layout == "nchw":
        B, Dd, Hc, Wc = desc.shape
        sB, sD, sH, sW = desc.stride()
    else:
        B, Hc, Wc, Dd = desc.shape
        sB, sH, sW, sD = desc.stride()
    D = Dd if D is None else D
    N = pts.shape[1]
    if out is None:
        out = torch.zeros((B, N, D), dtype=torch.float32, device=desc.device)
    _lib.check(L.yp_sample_desc(desc.data_ptr(), B, D, Hc, Wc, sB, sD, sH, sW, int(img_hw[0]), int(img_hw[1]), pts.data_ptr(),
                                count.data_ptr() if count is not None else None, N, out.data_ptr(), _stream(desc.device)))
    return out


def match_partial(d1: torch.Tensor, n1: Optional[torch.Tensor], d2: torch.Tensor, n2: Optional[torch.Tensor], col_off: int = 0,
                  out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """d1 [N1,D], d2 [N2,D] fp32 rows -> (row_key int64 [N1], col_key int64 [N2]); keys = dist_bits<<32 | index."""
    _need_cuda(d1, "desc1")
    L = _lib.lib(require_device=True)
    assert d1.is_contiguous() and d2.is_contiguous() and d1.dtype == torch.float32 and d2.dtype == torch.float32
    N1, D = d1.shape
    N2 = d2.shape[0]
    dev = d1.device
    if out is None:
        out = (torch.empty((N1,), dtype=torch.int64, device=dev), torch.empty((N2,), dtype=torch.int64, device=dev))
    rk, ck = out
    _lib.check(L.yp_match_partial(d1.data_ptr(), n1.data_ptr() if n1 is not None else None, N1, d2.data_ptr(),
                                  n2.data_ptr() if n2 is not None else None, N2, D, int(col_off), rk.data_ptr(), ck.data_ptr(), _stream(dev)))
    return rk, ck


def match_finalize(row_key: torch.Tensor, n1: Optional[torch.Tensor], col_key: torch.Tensor, nn_thresh: float,
                   out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """-> (matches [N1,3] fp32 rows (i, j, dist), count int32 [1])."""
    L = _lib.lib(require_device=True)
    N1, N2 = row_key.shape[0], col_key.shape[0]
    dev = row_key.device
    if out is None:
        out = (torch.zeros((N1, 3), dtype=torch.float32, device=dev), torch.zeros((1,), dtype=torch.int32, device=dev))
    m, cnt = out
    _lib.check(L.yp_match_finalize(row_key.data_ptr(), n1.data_ptr() if n1 is not None else None, N1, col_key.data_ptr(), N2,
                                   float(nn_thresh), m.data_ptr(), cnt.data_ptr(), _stream(dev)))
    return m, cnt


def _rowmin_pass(q: torch.Tensor, nq: Optional[torch.Tensor], k: torch.Tensor, nk: Optional[torch.Tensor], keys: torch.Tensor, col_off: int = 0,
                 col_keys: Optional[torch.Tensor] = None):
    """keys[i] = min_j key(q_i, k_j) on the tcgen05 conv kernel (YP_EPI_ROWMIN): q [Nq,D] are the 'pixels' of a 1x1 conv whose
    weights are the k [Nk,D] descriptors; 3xTF32 operands (hi/lo planes), the Nq x Nk similarity matrix is never written."""
    from ._lib import YP_ACT_NONE, YP_ALGO_TCGEN05, YP_EPI_ROWMIN, YP_FMT_F32X2, YpConvDesc
    from .engine import make_view, split_tf32
    L = _lib.lib(require_device=True)
    Nq, D = q.shape
    Nk = k.shape[0]
    Nkp = (Nk + 127) // 128 * 128                     # N tile of 128 output channels
    st = _stream(q.device)
    qa = torch.empty((2, 1, 1, Nq, D), dtype=torch.float32, device=q.device)
    _lib.check(L.yp_split_tf32(q.data_ptr(), Nq * D, qa[0].data_ptr(), qa[1].data_ptr(), st))
    kw = torch.empty((2, Nkp, D), dtype=torch.float32, device=k.device)
    if Nkp > Nk:
        kw[:, Nk:].zero_()                            # padding rows of the last N tile
    _lib.check(L.yp_split_tf32(k.data_ptr(), Nk * D, kw[0].data_ptr(), kw[1].data_ptr(), st))
    d = YpConvDesc()
    d.in_ = make_view(qa, YP_FMT_F32X2, 0, D)
    d.weight, d.bias = kw.data_ptr(), None
    d.ksize, d.stride, d.cout, d.act, d.epilogue, d.n_out = 1, 1, Nkp, YP_ACT_NONE, YP_EPI_ROWMIN, 0
    d.algo, d.tile_n, d.split_k = YP_ALGO_TCGEN05, 128, 1
    d.row_key = keys.data_ptr()
    d.col_key = col_keys.data_ptr() if col_keys is not None else None      # column minima from the same tiles (one pass for both directions)
    d.n_rows = nq.data_ptr() if nq is not None else None
    if nk is None:
        nk = _const_i32(k.device, Nk)                 # cached device scalar (a fresh torch.tensor would be a blocking H2D copy per call)
    d.n_cols = nk.data_ptr()
    d.col_off = int(col_off)
    _lib.check(L.yp_conv2d_nhwc_fwd(C.byref(d), _stream(q.device)))
    return qa, kw, nk      # keep the operand buffers alive until the caller has synchronised / enqueued its consumer


def match_partial_tc(d1: torch.Tensor, n1: Optional[torch.Tensor], d2: torch.Tensor, n2: Optional[torch.Tensor], col_off: int = 0):
    """Tensor-core form of match_partial: ONE pass over the N1 x N2 similarity tiles; the epilogue reduces the row minima (best d2 for
    every d1) and, through a shared-memory transpose of the tile, the column minima (best d1 for every d2).  ``YP_MATCH_TWO_PASS=1``
    selects the earlier form (a second row-minimum pass over the transposed problem)."""
    _need_cuda(d1, "desc1")
    assert d1.is_contiguous() and d2.is_contiguous() and d1.dtype == torch.float32 and d2.dtype == torch.float32
    assert d1.shape[1] % 16 == 0, "descriptor dimension must be a multiple of 16"
    rk = torch.full((d1.shape[0],), -1, dtype=torch.int64, device=d1.device)      # all ones = "no candidate"
    ck = torch.full((d2.shape[0],), -1, dtype=torch.int64, device=d1.device)
    import os
    if os.environ.get("YP_MATCH_TWO_PASS", "0") != "0":
        keep = [_rowmin_pass(d1, n1, d2, n2, rk, col_off), _rowmin_pass(d2, n2, d1, n1, ck, 0)]
    else:
        keep = [_rowmin_pass(d1, n1, d2, n2, rk, col_off, col_keys=ck)]
    rk._yp_keep = keep
    return rk, ck


def match_two_way(d1: torch.Tensor, n1, d2: torch.Tensor, n2, nn_thresh: float, algo: str = "auto"):
    """Single-GPU two-way match of row-major descriptors.  algo: "simt" = fused fp32 FMA kernel (yp_match_partial), "tc" = one
    3xTF32 tcgen05 pass over the similarity tiles whose epilogue reduces row AND column minima (yp_conv2d_nhwc_fwd + YP_EPI_ROWMIN with
    col_key), "auto" = tc from 4096 descriptors per side on (measured on B200, D = 256: 4096: 0.153 vs 0.285 ms, 8192: 0.43 vs
    1.12 ms, 16384: 1.46 vs 4.38 ms; at 2048: 0.166 vs 0.092 ms -- the fixed cost of the pass, ~0.14 ms, exceeds what it saves)."""
    if nn_thresh < 0.0:
        raise ValueError("'nn_thresh' should be non-negative")
    if algo == "auto":
        algo = "tc" if min(d1.shape[0], d2.shape[0]) >= 4096 and d1.shape[1] % 16 == 0 else "simt"
    rk, ck = match_partial_tc(d1, n1, d2, n2) if algo == "tc" else match_partial(d1, n1, d2, n2)
    return match_finalize(rk, n1, ck, nn_thresh)


class TwoWayMatcher:
    """match_two_way for descriptor sets of fixed capacity as ONE CUDA graph over static buffers: key initialisation, operand
    split, similarity pass (or the SIMT kernel), mutual check replay from one graph launch, which removes the ~60 us of per-call
    host work (seven launches + allocations) that dominates below ~8192 descriptors.  ``n1`` / ``n2`` are device counts (int32 [1])
    read by the kernels at replay time, so one matcher serves every frame of a stream; results live in ``self.matches`` /
    ``self.count`` until the next call."""

    def __init__(self, n1_cap: int, n2_cap: int, D: int, device, nn_thresh: float, algo: str = "auto"):
        dev = torch.device(device)
        self.d1 = torch.zeros((n1_cap, D), dtype=torch.float32, device=dev)
        self.d2 = torch.zeros((n2_cap, D), dtype=torch.float32, device=dev)
        self.n1 = torch.full((1,), n1_cap, dtype=torch.int32, device=dev)
        self.n2 = torch.full((1,), n2_cap, dtype=torch.int32, device=dev)
        if algo == "auto":      # without the per-call host work the tensor-core pass wins from 1024 descriptors on (0.030 vs 0.038 ms; 2048: 0.046 vs 0.088)
            algo = "tc" if min(n1_cap, n2_cap) >= 1024 and D % 16 == 0 else "simt"
        self.nn_thresh, self.algo = float(nn_thresh), algo
        self.graph = None
        self.matches = self.count = None

    def __call__(self, d1: torch.Tensor, d2: torch.Tensor, n1: Optional[int] = None, n2: Optional[int] = None):
        """d1 [<= n1_cap, D], d2 [<= n2_cap, D] device tensors (copied into the static operands) -> (matches [n1_cap,3], count [1])."""
        r1, r2 = d1.shape[0], d2.shape[0]
        if d1.data_ptr() != self.d1.data_ptr():
            self.d1[:r1].copy_(d1)
        if d2.data_ptr() != self.d2.data_ptr():
            self.d2[:r2].copy_(d2)
        self.n1.fill_(r1 if n1 is None else int(n1))
        self.n2.fill_(r2 if n2 is None else int(n2))
        if self.graph is None:
            run = lambda: match_two_way(self.d1, self.n1, self.d2, self.n2, self.nn_thresh, algo=self.algo)
            run()                                           # warm-up: fills the constant / workspace caches, sets kernel attributes
            torch.cuda.synchronize(self.d1.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.matches, self.count = run()
            self.graph = g
        self.graph.replay()
        return self.matches, self.count


def _linspace_pair(H: int, W: int, dev):
    """torch.linspace(-1, 1, n) for the columns / rows, exactly the values warp_image_batch builds its grid from."""
    return torch.linspace(-1, 1, W).to(dev), torch.linspace(-1, 1, H).to(dev)


def warp_batch(img: torch.Tensor, hinv: torch.Tensor, mode: str = "bilinear", out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """img [B,C,H,W] fp32, hinv [B,3,3] -> [B,C,H,W] sampled at the inverse-homography position of every output pixel
    (utils/utils.py:333-376: align_corners=True, zeros padding)."""
    _need_cuda(img, "img")
    L = _lib.lib(require_device=True)
    assert mode in ("bilinear", "nearest"), mode
    img = img.contiguous().float()
    B, Cc, H, W = img.shape
    hinv = hinv.to(img.device).float().contiguous().view(-1, 3, 3)
    assert hinv.shape[0] == B, (hinv.shape, B)
    xs, ys = _linspace_pair(H, W, img.device)
    if out is None:
        out = torch.empty_like(img)
    _lib.check(L.yp_warp_image_batch(img.data_ptr(), hinv.data_ptr(), xs.data_ptr(), ys.data_ptr(), B, Cc, H, W, int(mode == "nearest"),
                                     out.data_ptr(), _stream(img.device)))
    return out


def homography_adapt(heat: torch.Tensor, mask: torch.Tensor, hinv: torch.Tensor, want_sums: bool = False):
    """heat, mask [B,H,W] fp32 (B warped copies of one image, their valid masks), hinv [B,3,3] -> aggregated heatmap [H,W]
    = sum_b warp(heat_b * mask_b) / sum_b warp(mask_b) in one kernel (export_homography.py:97-128)."""
    _need_cuda(heat, "heatmap")
    L = _lib.lib(require_device=True)
    heat, mask = heat.contiguous().float(), mask.to(heat.device).contiguous().float()
    B, H, W = heat.shape
    assert mask.shape == heat.shape
    hinv = hinv.to(heat.device).float().contiguous().view(-1, 3, 3)
    assert hinv.shape[0] == B
    xs, ys = _linspace_pair(H, W, heat.device)
    agg = torch.empty((H, W), dtype=torch.float32, device=heat.device)
    sh = torch.empty_like(agg) if want_sums else None
    sm = torch.empty_like(agg) if want_sums else None
    _lib.check(L.yp_homography_adaptation(heat.data_ptr(), mask.data_ptr(), hinv.data_ptr(), xs.data_ptr(), ys.data_ptr(), B, H, W,
                                          sh.data_ptr() if want_sums else None, sm.data_ptr() if want_sums else None, agg.data_ptr(),
                                          _stream(heat.device)))
    return (agg, sh, sm) if want_sums else agg
