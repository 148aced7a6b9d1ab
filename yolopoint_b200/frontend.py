"""Whole-frame pipeline: the body of ``YoloPointFrontend.process_img`` (src/demo.py:125-230 of the reference) plus the
tracker's two-way match against the previous frame (src/demo.py:386, 300-341), as ONE CUDA graph per frame parity:

  uint8 frame -> space-to-depth operand -> network -> Detect decode -> box NMS -> cell softmax heatmap ->
  keypoint NMS (+ border and in-box filters) -> descriptor sampling -> match with the previous frame

The reference crosses the host/device boundary four times per frame and runs NMS / matching on the CPU; here the
frame goes up once and the compact results come back once.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import YpNmsParams
from .engine import SliceRef

DEFAULT_CFG = dict(  # configs/kitti_inference.yaml:5-16 of the reference
    detection_threshold=0.12, nms=8, nn_thresh=0.7, conf_thres_box=0.4, iou_thres_box=0.45, max_det=1000,
)


class FramePipeline:
    """Static buffers + graphs for frames of one shape.  All results stay on the device until ``fetch``."""

    def __init__(self, model, B: int, H: int, W: int, cfg: Optional[dict] = None, filter_pts: bool = True, max_pts: int = 4096,
                 nms_cap: int = 4096, heat_variant: int = 1, do_match: bool = True, slot: int = 0):
        self.cfg = dict(DEFAULT_CFG, **(cfg or {}))
        self.eng = model.engine() if hasattr(model, "engine") else model
        self.plan = self.eng.plan(B, H, W, slot)
        self.B, self.H, self.W = B, H, W
        self.filter_pts, self.max_pts, self.nms_cap, self.heat_variant, self.do_match = filter_pts, max_pts, (nms_cap + 63) // 64 * 64, heat_variant, do_match
        dev, D = self.eng.device, self.eng.net.D
        self.D = D
        L = _lib.lib(require_device=True)
        md = self.cfg["max_det"]
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)
        self.boxes, self.bcount = z(B, md, 6), z(B, dt=torch.int32)
        self.heat = z(B, H, W)
        self.pts = [z(B, max_pts, 3), z(B, max_pts, 3)]
        self.kcount = [z(B, dt=torch.int32), z(B, dt=torch.int32)]
        self.descs = [z(B, max_pts, D), z(B, max_pts, D)]
        self.row_key, self.col_key = z(B, max_pts, dt=torch.int64), z(B, max_pts, dt=torch.int64)
        self.matches, self.mcount = z(B, max_pts, 3), z(B, dt=torch.int32)
        self.ws_nms = torch.empty(L.yp_box_nms_workspace_bytes(B, self.plan.A, self.eng.net.no, self.nms_cap), dtype=torch.uint8, device=dev)
        self.ws_kp = torch.empty(L.yp_keypoints_workspace_bytes(B, H, W, max_pts), dtype=torch.uint8, device=dev)
        self.nms_params = YpNmsParams(float(self.cfg["conf_thres_box"]), float(self.cfg["iou_thres_box"]), 1, 1, int(md), 30000, 7680.0, None)
        self.parity = 0
        # pinned host mirrors for the single read-back, double-buffered so that one frame can be staged / unpacked on the host
        # while the previous one is still on the GPU (submit_host may run one frame ahead of collect)
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        self.d_counts = z(3, B, dt=torch.int32)
        self._host = [dict(frame=torch.empty((B, H, W, 3), dtype=torch.uint8, pin_memory=True),
                           counts=torch.empty((3, B), dtype=torch.int32, pin_memory=True), pts=pin(self.pts[0]), boxes=pin(self.boxes),
                           desc=pin(self.descs[0]), matches=pin(self.matches), done=None) for _ in range(2)]
        self.h_frame, self.h_counts = self._host[0]["frame"], self._host[0]["counts"]
        self.h_pts, self.h_boxes, self.h_desc, self.h_matches = (self._host[0][k] for k in ("pts", "boxes", "desc", "matches"))
        self._n_submit = self._n_collect = 0
        self.graphs = {}   # graphs bake THIS pipeline's buffer addresses, so they are owned here, not by the shared plan

    # ---- device work ---------------------------------------------------------------------------
    def _enqueue(self, k: int, from_frame: bool = True):
        """Everything for the frame currently in plan.frame_in (or plan.x_in), results into parity-k buffers."""
        L, p, dev = _lib.lib(), self.plan, self.eng.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        B, H, W, cfg = self.B, self.H, self.W, self.cfg
        p.run_input(from_frame)
        semi = p.bufs["semi"][0]   # [B,Hc,Wc,80] fp32 NHWC
        sB, sH, sW, sC = semi.stride()

        def kp_tail(stp):   # runs on the keypoint head's stream, overlapping the detection branch
            _lib.check(L.yp_heatmap(semi.data_ptr(), B, H // 8, W // 8, sB, sC, sH, sW, self.heat_variant, self.heat.data_ptr(), stp))
            _lib.check(L.yp_keypoints_nms(self.heat.data_ptr(), B, H, W, float(cfg["detection_threshold"]), int(cfg["nms"]), self.max_pts,
                                          self.ws_kp.data_ptr(), self.ws_kp.numel(), stp))

        p.run_net(tails={1: kp_tail})
        # Detect decode fused into the NMS front end: pred [B,A,85] is never materialised here
        dets = [p.bufs[f"det{i}"] for i in range(3)]
        lg = (C.c_void_p * 3)(*[d.data_ptr() for d in dets])
        ny = (C.c_int32 * 3)(*[d.shape[2] for d in dets]); nx = (C.c_int32 * 3)(*[d.shape[3] for d in dets])
        ldc = (C.c_int32 * 3)(*[d.shape[4] for d in dets])
        strd = (C.c_float * 3)(*[float(v) for v in self.eng.stride])
        anc = (C.c_float * 18)(*[float(v) for row in self.eng.anchors_px for v in row])
        _lib.check(L.yp_detect_nms(lg, ny, nx, ldc, strd, anc, B, 3, self.eng.net.no, C.byref(self.nms_params), self.nms_cap,
                                   self.boxes.data_ptr(), self.bcount.data_ptr(), self.ws_nms.data_ptr(), self.ws_nms.numel(), st))
        _lib.check(L.yp_keypoints_collect(self.heat.data_ptr(), B, H, W, 4, self.boxes.data_ptr() if self.filter_pts else None,
                                          self.bcount.data_ptr() if self.filter_pts else None, self.boxes.shape[1] if self.filter_pts else 0,
                                          self.pts[k].data_ptr(), self.kcount[k].data_ptr(), self.max_pts, self.ws_kp.data_ptr(),
                                          self.ws_kp.numel(), st))
        desc = p.bufs["desc"][0]   # [B,Hc,Wc,D] fp32 NHWC, unit norm
        dB, dH, dW, dD = desc.stride()
        _lib.check(L.yp_sample_desc(desc.data_ptr(), B, self.D, H // 8, W // 8, dB, dD, dH, dW, H, W, self.pts[k].data_ptr(),
                                    self.kcount[k].data_ptr(), self.max_pts, self.descs[k].data_ptr(), st))
        if self.do_match:
            for b in range(B):  # previous frame (parity 1-k) is desc1, current is desc2, as PointTracker.update does
                _lib.check(L.yp_match_partial(self.descs[1 - k][b].data_ptr(), self.kcount[1 - k][b:].data_ptr(), self.max_pts,
                                              self.descs[k][b].data_ptr(), self.kcount[k][b:].data_ptr(), self.max_pts, self.D, 0,
                                              self.row_key[b].data_ptr(), self.col_key[b].data_ptr(), st))
                _lib.check(L.yp_match_finalize(self.row_key[b].data_ptr(), self.kcount[1 - k][b:].data_ptr(), self.max_pts,
                                               self.col_key[b].data_ptr(), self.max_pts, float(cfg["nn_thresh"]), self.matches[b].data_ptr(),
                                               self.mcount[b:].data_ptr(), st))
        self.d_counts[0].copy_(self.kcount[k]); self.d_counts[1].copy_(self.bcount); self.d_counts[2].copy_(self.mcount)

    def n_launches(self) -> int:
        """Kernels of this library launched per frame batch (for bench.py's gpu_launches)."""
        net_launches = len(self.plan.launches)
        # input + net + fused decode/NMS (candidates, counts, rank, mask, scan) + heatmap + keypoints (nms, collect, emit) + sample + match
        # keypoints = 8 NMS rounds + 1 sweep + collect + emit (see csrc/keypoints.cu)
        per = 1 + net_launches + 5 + 1 + 11 + 1 + (self.B * 3 if self.do_match else 0)
        return per

    def step_device(self, from_frame: bool = True):
        """Process the frame already resident in plan.frame_in / plan.x_in; flips the parity."""
        k = self.parity
        key = (k, bool(from_frame))
        g = self.graphs.get(key)
        if g is None:
            self._enqueue(k, from_frame)            # eager first run: sets function attributes, fills caches
            torch.cuda.synchronize(self.eng.device)
            if self.eng.use_graphs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(k, from_frame)
                self.graphs[key] = g
                # the eager run already produced this frame's results; replaying is idempotent (same inputs)
        else:
            g.replay()
        self.parity = 1 - k
        return k

    def reset_tracking(self):
        for c in self.kcount:
            c.zero_()
        self.parity = 0

    # ---- host boundary -------------------------------------------------------------------------
    def submit_host(self, frames_u8: np.ndarray):
        """Enqueue (on the current stream) H2D of the frames, the whole pipeline and the D2H of the compact results.  At most two
        submissions may be outstanding (double-buffered pinned host memory): ``submit(i+1)`` before ``collect(i)`` lets the host
        stage / unpack one frame while the GPU works on the other."""
        if self._n_submit - self._n_collect >= 2:
            raise RuntimeError("FramePipeline: two frames already in flight; call collect() first")
        dev = self.eng.device
        h = self._host[self._n_submit % 2]
        h["frame"].copy_(torch.from_numpy(np.ascontiguousarray(frames_u8)).view(self.B, self.H, self.W, 3))
        self.plan.frame_in.copy_(h["frame"], non_blocking=True)
        k = self.step_device(True)
        h["counts"].copy_(self.d_counts, non_blocking=True)
        h["pts"].copy_(self.pts[k], non_blocking=True)
        h["boxes"].copy_(self.boxes, non_blocking=True)
        h["desc"].copy_(self.descs[k], non_blocking=True)
        h["matches"].copy_(self.matches, non_blocking=True)
        h["done"] = torch.cuda.Event()
        h["done"].record(torch.cuda.current_stream(dev))
        self._n_submit += 1

    def collect(self):
        """Wait for the oldest outstanding submit_host and unpack per-image (pts[3,N] f64, desc[D,N] f32, boxes[n,6] f32,
        matches[3,L] f64)."""
        if self._n_collect >= self._n_submit:
            raise RuntimeError("FramePipeline.collect() without a matching submit_host()")
        h = self._host[self._n_collect % 2]
        self._n_collect += 1
        h["done"].synchronize()
        out = []
        for b in range(self.B):
            nk, nb, nm = (int(v) for v in h["counts"][:, b])
            if nk < 0 or nb < 0:
                raise RuntimeError(f"buffer overflow (keypoints {nk}, boxes {nb}): raise max_pts / nms_cap")
            pts = h["pts"][b, :nk].numpy().astype(np.float64).T.copy()
            desc = h["desc"][b, :nk].numpy().T.copy()
            boxes = h["boxes"][b, :nb].numpy().copy()
            matches = h["matches"][b, :max(nm, 0)].numpy().astype(np.float64).T.copy()
            out.append((pts, desc, boxes, matches))
        return out

    def step_host(self, frames_u8: np.ndarray):
        """frames [B,H,W,3] uint8 on the host -> per-image (pts, desc, boxes, matches); one H2D, one D2H, one sync."""
        self.submit_host(frames_u8)
        return self.collect()

    def d2h_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.h_counts, self.h_pts, self.h_boxes, self.h_desc, self.h_matches))

    def h2d_bytes(self) -> int:
        return self.h_frame.numel()


class YoloPointFrontend:
    """Drop-in for the reference frontend's per-frame call (src/demo.py:15-230) around an already-built model.

    ``process_img(img, rostpc=None)`` takes a uint8 HxWx3 frame and returns ``(pts [3,N] float64, desc [D,N] float32, [boxes [n,6]])``
    exactly like the reference: crop to multiples of 32, the optional ``crop_resize`` window of the config (crop + ``cv2.resize`` on the
    host, as the reference does it) and the optional per-camera template mask (``self.templates[rostpc]``, src/demo.py:187-192) included.
    """

    def __init__(self, model, config: Optional[dict] = None, filter_pts: bool = True, max_pts: int = 4096, nms_cap: int = 4096):
        self.model = model
        cfg = dict(DEFAULT_CFG)
        self.crop_resize = None
        if config:
            m = config.get("model", config)
            cfg.update({k: v for k, v in m.get("superpoint", {}).items() if k in cfg})
            cfg.update({k: v for k, v in m.get("yolo", {}).items() if k in cfg})
            cfg.update({k: v for k, v in config.items() if k in cfg})
            self.crop_resize = config.get("crop_resize")        # src/demo.py:52
        self.cfg, self.filter_pts, self.max_pts, self.nms_cap = cfg, filter_pts, max_pts, nms_cap
        self.cell, self.border_remove = 8, 4
        self.templates: Dict[str, np.ndarray] = {}              # camera name -> [H,W] mask, 1 = keep (src/demo.py:94)
        self._pipes: Dict[Tuple[int, int], FramePipeline] = {}
        self.last_matches = None

    def preprocess(self, img, interpolation=None):
        """src/demo.py:97-123: optional crop + resize to width ``crop_resize[4]``, then a centred crop to multiples of 32 (decided,
        like the reference, on the shape of the frame as it came in).  -> (img, cut_h0, cut_w0, resize_fac)"""
        shape0 = img.shape[:2]
        resize_fac = 1.0
        cut_h0 = cut_w0 = 0
        if self.crop_resize:
            import cv2
            cr = self.crop_resize
            w1 = cr[4]
            img = img[cr[0]:cr[1], cr[2]:cr[3]]
            h0, w0 = img.shape[:2]
            resize_fac = w1 / w0
            img = cv2.resize(img, (w1, round(h0 * resize_fac)), interpolation=cv2.INTER_LINEAR if interpolation is None else interpolation)
        if shape0[0] % 32 or shape0[1] % 32:
            h0, w0 = img.shape[:2]
            cut_h, cut_w = (h0 % 32) / 2, (w0 % 32) / 2
            cut_h0, cut_h1 = int(np.ceil(cut_h)), int(np.floor(cut_h))
            cut_w0, cut_w1 = int(np.ceil(cut_w)), int(np.floor(cut_w))
            img = img[cut_h0:h0 - cut_h1, cut_w0:w0 - cut_w1]
        return img, cut_h0, cut_w0, resize_fac

    def template_filter(self, pts: np.ndarray, desc: np.ndarray, rostpc):
        """Second half of the reference's ``filter_points`` closure (src/demo.py:187-195): keep the points whose pixel of the camera's
        template mask equals 1.  Descriptor sampling is per point, so dropping columns afterwards equals filtering before sampling."""
        try:
            template = self.templates[rostpc]
        except KeyError:
            print(f"Template for {rostpc} not found.")
            return pts, desc, False
        keep = (np.ones(template.shape[:2]) * template)[pts[1].astype(int), pts[0].astype(int)] == 1
        if keep.all():
            return pts, desc, False
        return pts[:, keep], (desc[:, keep] if keep.any() else np.zeros((desc.shape[0], 0))), True

    def restore_coords(self, pts: np.ndarray, boxes: torch.Tensor, cth: int, ctw: int, resize_fac: float):
        """src/demo.py:217-228: back to the coordinates of the frame as it came in.  With ``crop_resize`` the reference then adds the
        crop offsets through ``pts[:, 0] += cr[2]; pts[:, 1] += cr[0]`` on the [3,N] array, i.e. to the first two POINTS (all three
        rows) rather than to the x / y rows; reproduced as is so that results stay identical."""
        pts[0] = (pts[0] + ctw) / resize_fac
        pts[1] = (pts[1] + cth) / resize_fac
        boxes[:, :4] = (boxes[:, :4] + torch.tensor([ctw, cth, ctw, cth], dtype=boxes.dtype)) / resize_fac
        if self.crop_resize:
            cr = self.crop_resize
            pts[:, 0] += cr[2]
            pts[:, 1] += cr[0]
            boxes[:, :4] += torch.tensor([cr[2], cr[0], cr[2], cr[0]], dtype=boxes.dtype)
        return pts, boxes

    def pipeline(self, H: int, W: int) -> FramePipeline:
        if (H, W) not in self._pipes:
            self._pipes[(H, W)] = FramePipeline(self.model, 1, H, W, self.cfg, self.filter_pts, self.max_pts, self.nms_cap)
        return self._pipes[(H, W)]

    @torch.no_grad()
    def process_img(self, img, rostpc=None):
        img, cth, ctw, resize_fac = self.preprocess(img)
        H, W, _ = img.shape
        pts, desc, boxes, matches = self.pipeline(H, W).step_host(img[None])[0]
        self.last_matches = matches
        obj = torch.from_numpy(boxes)
        if pts.shape[1] == 0:
            return np.zeros((3, 0)), None, None  # src/demo.py:152-153
        if self.filter_pts and rostpc:
            pts, desc, dropped = self.template_filter(pts, desc, rostpc)
            if dropped:
                self.last_matches = None          # the device-side matches index the unfiltered sets
        pts, obj = self.restore_coords(pts, obj, cth, ctw, resize_fac)
        return pts, desc, [obj]
