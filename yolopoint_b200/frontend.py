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

NMS_CAP = 30016     # >= the reference's max_nms = 30000 (src/utils/general_yolo.py:155), multiple of 64

DEFAULT_CFG = dict(  # configs/kitti_inference.yaml:5-16 of the reference
    detection_threshold=0.12, nms=8, nn_thresh=0.7, conf_thres_box=0.4, iou_thres_box=0.45, max_det=1000,
)


class FramePipeline:
    """Static buffers + graphs for frames of one shape.  All results stay on the device until ``fetch``."""

    def __init__(self, model, B: int, H: int, W: int, cfg: Optional[dict] = None, filter_pts: bool = True, max_pts: Optional[int] = None,
                 nms_cap: int = NMS_CAP, heat_variant: int = 1, do_match: bool = True, slot: int = 0, frames_in_flight: int = 1):
        """``max_pts`` / ``nms_cap`` are buffer capacities, not semantics.  The defaults cannot overflow: ``max_pts=None`` is the
        geometric bound of the keypoint NMS (survivors are more than ``nms`` pixels apart) and ``nms_cap`` = 30016 covers the
        reference's ``max_nms`` = 30000 candidates, beyond which the kernel keeps the 30000 most confident like the reference
        (src/utils/general_yolo.py:155, 210-211).  Smaller explicit values save memory; a frame that exceeds them makes the
        pipeline grow to the defaults and re-run the frame -- valid input never raises.

        ``frames_in_flight`` = F > 1 software-pipelines consecutive frames of the ONE camera stream: every frame is still a batch-1 pass
        with the same kernels and bit-identical results, but frame i+1's network starts while frame i's network tail / NMS / match are
        still running (at batch 1 a layer occupies 4-200 CTAs of a 148-SM GPU, so two or three frames fill it better than one).  Only the
        in-box filter + match of frame i+1 wait for frame i (they read its keypoint count and descriptors, src/demo.py:386).  Every
        per-frame buffer then exists F + 1 times (frame i + F + 1 reuses the buffers of frame i once frame i+1 has matched against them),
        every context has its own activation buffers, stream and CUDA graphs."""
        self.cfg = dict(DEFAULT_CFG, **(cfg or {}))
        self.eng = model.engine() if hasattr(model, "engine") else model
        self.F = max(1, int(frames_in_flight))
        self.nctx = 2 if self.F == 1 else self.F + 1          # F == 1: two result parities on one stream (frame i+1 is computed while frame i is read back)
        self.plans = [self.eng.plan(B, H, W, slot)] * 2 if self.F == 1 else [self.eng.plan(B, H, W, 16 * slot + c) for c in range(self.nctx)]
        self.B, self.H, self.W = B, H, W
        self.filter_pts, self.heat_variant, self.do_match = filter_pts, heat_variant, do_match
        self.max_pts = min(int(max_pts), self.max_pts_bound()) if max_pts else self.max_pts_bound()
        self.nms_cap = (min(int(nms_cap), NMS_CAP) + 63) // 64 * 64
        self.D = self.eng.net.D
        self.parity = 0
        self._n_submit = self._n_collect = 0
        self.regrown = 0      # how many times a frame exceeded an explicit capacity (diagnostics)
        self._alloc()

    @property
    def plan(self):
        """Activation buffers / launch list of the context the NEXT frame runs in (``plan.frame_in`` / ``plan.x_in`` is where that frame
        goes).  With frames in flight the context's previous frame may still be queued on its own stream: the current stream is made
        to wait for it here, so that whatever the caller enqueues next on the current stream (the copy of the new frame into
        ``frame_in``) cannot overtake the input conversion of the frame that used the buffer before."""
        k = self.parity
        if self.F > 1 and self._ctx_used[k]:
            torch.cuda.current_stream(self.eng.device).wait_event(self.ev_frame[k])
        return self.plans[k]

    def max_pts_bound(self) -> int:
        """Keypoints that can survive nms_fast with radius r on an H x W frame: survivors are >= r+1 pixels apart (Chebyshev)."""
        r = int(self.cfg["nms"]) + 1
        return (-(-self.H // r) * -(-self.W // r) + 63) // 64 * 64

    def _alloc(self, old: Optional[dict] = None):
        """(Re)allocate every capacity-dependent buffer; ``old`` carries the previous frames' keypoints / descriptors over."""
        B, H, W, D = self.B, self.H, self.W, self.D
        max_pts = self.max_pts
        dev = self.eng.device
        L = _lib.lib(require_device=True)
        md = self.cfg["max_det"]
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)
        n = self.nctx
        two = lambda f: [f() for _ in range(n)]
        # scratch of one frame: shared by both parities when frames run one after the other, one per context when they overlap
        def scratch(f):
            if self.F == 1:
                x = f()
                return [x, x]
            return [f() for _ in range(n)]
        # every result buffer exists once per frame parity / context: frame i+1 is computed while frame i's results are read back
        self.boxes, self.bcount = two(lambda: z(B, md, 6)), two(lambda: z(B, dt=torch.int32))
        self.heat = scratch(lambda: z(B, H, W))
        self.pts, self.kcount = two(lambda: z(B, max_pts, 3)), two(lambda: z(B, dt=torch.int32))
        self.descs = two(lambda: z(B, max_pts, D))                       # box-filtered descriptors, compact (what the host reads)
        # NMS survivors inside the border BEFORE the in-box filter and their descriptors: built while the detection branch is still
        # running; `sel` = indices of the points that passed the filter.  The match reads rows sel[i] of descs_all directly.
        self.pts_all, self.n_all = scratch(lambda: z(B, max_pts, 3)), scratch(lambda: z(B, dt=torch.int32))
        self.descs_all, self.sel = two(lambda: z(B, max_pts, D)), two(lambda: z(B, max_pts, dt=torch.int32))
        self.row_key, self.col_key = scratch(lambda: z(B, max_pts, dt=torch.int64)), scratch(lambda: z(B, max_pts, dt=torch.int64))
        self.matches, self.mcount = two(lambda: z(B, max_pts, 3)), two(lambda: z(B, dt=torch.int32))
        self.d_counts = two(lambda: z(4, B, dt=torch.int32))            # keypoints, boxes, matches, pixels >= detection threshold
        self.ws_nms = scratch(lambda: torch.zeros(L.yp_box_nms_workspace_bytes(B, self.plans[0].A, self.eng.net.no, self.nms_cap), dtype=torch.uint8, device=dev))
        self.ws_kp = scratch(lambda: torch.empty(L.yp_keypoints_workspace_bytes(B, H, W, max_pts), dtype=torch.uint8, device=dev))
        self.nms_params = YpNmsParams(float(self.cfg["conf_thres_box"]), float(self.cfg["iou_thres_box"]), 1, 1, int(md), 30000, 7680.0, None)
        # host boundary: frames go up on a copy stream into a device staging buffer, results come back on another copy stream
        # (counts first, then exactly `count` rows of each result), so that neither transfer sits between two frames' kernels
        if old is None:
            # three copy streams: frames up, counts down, result rows down.  The rows of frame i are requested (collect) after the
            # counts of frame i+1 were enqueued (submit), so they need their own stream -- on a shared in-order stream they
            # would wait for frame i+1 to finish (measured: a 200 us hole between frames)
            self.s_in, self.s_out, self.s_res = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            # F > 1: one compute stream and one completion event per context
            self.cs = [torch.cuda.Stream(dev) for _ in range(n)] if self.F > 1 else []
            self.ev_frame = [torch.cuda.Event() for _ in range(n)]
            self._ctx_used = [False] * n
            self.d_frame = two(lambda: torch.zeros((B, H, W, 3), dtype=torch.uint8, device=dev))
            self._host = [dict(frame=torch.empty((B, H, W, 3), dtype=torch.uint8, pin_memory=True),
                               counts=torch.zeros((4, B), dtype=torch.int32, pin_memory=True), k=0, used=False,
                               ev_in=torch.cuda.Event(), ev_read=torch.cuda.Event(), ev_counts=torch.cuda.Event(), ev_done=torch.cuda.Event())
                          for _ in range(n)]
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        for h in self._host:
            h.update(pts=pin(self.pts[0]), boxes=pin(self.boxes[0]), desc=pin(self.descs[0]), matches=pin(self.matches[0]))
            h.update({n + "_np": h[n].numpy() for n in ("frame", "counts", "pts", "boxes", "desc", "matches")})
        self.graphs = {}   # graphs bake THIS pipeline's buffer addresses, so they are owned here, not by the shared plan
        self._d2h_bytes = 0
        if old is not None:   # the previous frame's results are the match partner of the next frame
            n_old = old["sel"][0].shape[1]
            for k in range(n):
                self.descs_all[k][:, :n_old].copy_(old["descs_all"][k]); self.sel[k][:, :n_old].copy_(old["sel"][k]); self.kcount[k].copy_(old["kcount"][k])

    # ---- device work ---------------------------------------------------------------------------
    def _enqueue(self, k: int, from_frame: bool = True, with_input: bool = True, part: int = 0):
        """Everything for the frame currently in plan.frame_in (or plan.x_in), results into the buffers of parity / context k.
        ``with_input=False``: the stem's operand buffer has already been filled (the host path converts straight from its staging
        buffer).  ``part`` 1 = everything that does not need the previous frame (input conversion, network, keypoint / descriptor
        tails, box NMS), 2 = in-box filter, match against context k-1, result counts; 0 = both."""
        L, p, dev = _lib.lib(), self.plans[k], self.eng.device
        prev = (k - 1) % self.nctx
        heat, pts_all, n_all, row_key, col_key, ws_kp, ws_nms = (self.heat[k], self.pts_all[k], self.n_all[k], self.row_key[k], self.col_key[k],
                                                                 self.ws_kp[k], self.ws_nms[k])
        if part == 2:
            return self._enqueue_tail(k, prev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        B, H, W, cfg = self.B, self.H, self.W, self.cfg
        if with_input:
            p.run_input(from_frame)
        semi = p.bufs["semi"][0]   # [B,Hc,Wc,80] fp32 NHWC
        sB, sH, sW, sC = semi.stride()

        desc = p.bufs["desc"][0]   # [B,Hc,Wc,D] fp32 NHWC, unit norm
        dB, dH, dW, dD = desc.stride()
        multi = self.eng.multi_stream   # False: every launch on the current stream, already ordered
        ev_kp = torch.cuda.Event() if multi else None
        nthr = self.d_counts[k][3]

        def kp_tail(stp):   # keypoint head's stream, under the detection branch: heatmap, NMS, border filter, confidence order
            _lib.check(L.yp_heatmap(semi.data_ptr(), B, H // 8, W // 8, sB, sC, sH, sW, self.heat_variant, heat.data_ptr(), stp))
            _lib.check(L.yp_keypoints_nms(heat.data_ptr(), B, H, W, float(cfg["detection_threshold"]), int(cfg["nms"]), self.max_pts,
                                          ws_kp.data_ptr(), ws_kp.numel(), stp))
            _lib.check(L.yp_keypoints_collect(heat.data_ptr(), B, H, W, 4, None, None, 0, pts_all.data_ptr(), n_all.data_ptr(),
                                              self.max_pts, ws_kp.data_ptr(), ws_kp.numel(), stp))
            _lib.check(L.yp_keypoints_threshold_count(ws_kp.data_ptr(), ws_kp.numel(), B, H, W, self.max_pts, nthr.data_ptr(), stp))
            if multi:
                ev_kp.record(p.side_stream(1))

        def desc_tail(stp):  # descriptor head's stream: sample every listed point (the in-box filter only drops points later)
            if multi:
                p.side_stream(2).wait_event(ev_kp)
            _lib.check(L.yp_sample_desc(desc.data_ptr(), B, self.D, H // 8, W // 8, dB, dD, dH, dW, H, W, pts_all.data_ptr(),
                                        n_all.data_ptr(), self.max_pts, self.descs_all[k].data_ptr(), stp))

        # Detect decode fused into the box NMS: pred [B,A,85] is never materialised here.  The candidates of levels 0 / 1 (95 % of
        # the rows) are listed on side streams as soon as their Detect convolution is done; the NMS kernel after the last layer
        # scans level 2 only and starts from that list.
        dets = [p.bufs[f"det{i}"] for i in range(3)]
        lg = (C.c_void_p * 3)(*[d.data_ptr() for d in dets])
        ny = (C.c_int32 * 3)(*[d.shape[2] for d in dets]); nx = (C.c_int32 * 3)(*[d.shape[3] for d in dets])
        ldc = (C.c_int32 * 3)(*[d.shape[4] for d in dets])
        strd = (C.c_float * 3)(*[float(v) for v in self.eng.stride])
        anc = (C.c_float * 18)(*[float(v) for row in self.eng.anchors_px for v in row])
        self._keep_alive = getattr(self, "_keep_alive", []) + [(lg, ny, nx, ldc, strd, anc)]

        def prescan(level):
            return lambda stp: _lib.check(L.yp_detect_prescan(lg, ny, nx, ldc, strd, anc, B, 3, self.eng.net.no, C.byref(self.nms_params),
                                                              self.nms_cap, level, ws_nms.data_ptr(), ws_nms.numel(), stp))

        p.run_net(tails={1: kp_tail, 2: desc_tail}, after={"Detect.m.0": prescan(0), "Detect.m.1": prescan(1)})
        boxes, bcount = self.boxes[k], self.bcount[k]
        _lib.check(L.yp_detect_nms(lg, ny, nx, ldc, strd, anc, B, 3, self.eng.net.no, C.byref(self.nms_params), self.nms_cap,
                                   boxes.data_ptr(), bcount.data_ptr(), ws_nms.data_ptr(), ws_nms.numel(), 3, st))
        if part == 0:
            self._enqueue_tail(k, prev)

    def _enqueue_tail(self, k: int, prev: int):
        """Critical path after the box NMS: in-box filter (order-preserving compaction) -> match with the previous frame (context
        ``prev``) -> result counts."""
        L, p, dev = _lib.lib(), self.plans[k], self.eng.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        B, H, W, cfg = self.B, self.H, self.W, self.cfg
        multi = self.eng.multi_stream
        boxes, bcount = self.boxes[k], self.bcount[k]
        pts_all, n_all, row_key, col_key = self.pts_all[k], self.n_all[k], self.row_key[k], self.col_key[k]
        fp = self.filter_pts
        _lib.check(L.yp_keypoints_filter(pts_all.data_ptr(), n_all.data_ptr(), B, self.max_pts, H, W, boxes.data_ptr() if fp else None,
                                         bcount.data_ptr() if fp else None, boxes.shape[1] if fp else 0, self.pts[k].data_ptr(),
                                         self.sel[k].data_ptr(), self.kcount[k].data_ptr(), self.kcount[prev].data_ptr(),
                                         row_key.data_ptr(), col_key.data_ptr(), st))
        main = torch.cuda.current_stream(dev)
        side = p.side_stream(2) if multi else main
        if multi:   # the compact descriptor block is only needed by the host: gather it beside the match
            ev_f = torch.cuda.Event(); ev_f.record(main); side.wait_event(ev_f)
        _lib.check(L.yp_gather_rows(self.descs_all[k].data_ptr(), self.sel[k].data_ptr(), self.kcount[k].data_ptr(), B, self.max_pts, self.D,
                                    self.descs[k].data_ptr(), C.c_void_p(side.cuda_stream)))
        dc = self.d_counts[k]
        if self.do_match:   # previous frame (parity 1-k) is desc1, current is desc2, as PointTracker.update does
            _lib.check(L.yp_match_frames(self.descs_all[prev].data_ptr(), self.sel[prev].data_ptr(), self.kcount[prev].data_ptr(),
                                         self.descs_all[k].data_ptr(), self.sel[k].data_ptr(), self.kcount[k].data_ptr(), B, self.max_pts,
                                         self.D, row_key.data_ptr(), col_key.data_ptr(), float(cfg["nn_thresh"]),
                                         self.matches[k].data_ptr(), self.mcount[k].data_ptr(), self.kcount[k].data_ptr(), bcount.data_ptr(),
                                         dc.data_ptr(), st))
        else:
            dc[0].copy_(self.kcount[k]); dc[1].copy_(bcount); dc[2].zero_()
        if multi:
            ev_g = torch.cuda.Event(); ev_g.record(side); main.wait_event(ev_g)

    def n_launches(self) -> int:
        """Kernels of this library launched per frame batch (for bench.py's gpu_launches)."""
        net_launches = self.plans[0].n_net_launches()
        # input + net + box-NMS candidate prescan of Detect levels 0 / 1 (2) + fused decode / box NMS (1) + heatmap + keypoints (8 NMS rounds + sweep + collect + emit) + sample + in-box filter +
        # descriptor gather + match (tiles + finalize, all images)
        return 1 + net_launches + 2 + 1 + 1 + 11 + 1 + 1 + 1 + (2 if self.do_match else 0)

    def _graphed(self, key, fn):
        """Replay the CUDA graph of ``fn`` (captured on first use after one eager run, which already produced the results)."""
        g = self.graphs.get(key)
        if g is None:
            fn()                                            # eager first run: sets function attributes, fills caches
            torch.cuda.synchronize(self.eng.device)
            if self.eng.use_graphs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                self.graphs[key] = g
                # replaying is idempotent (same inputs), so the eager results stand
        else:
            g.replay()

    def step_device(self, from_frame: bool = True, frame_src: Optional[torch.Tensor] = None):
        """Process the frame already resident in plan.frame_in / plan.x_in (or in ``frame_src``, a uint8 [B,H,W,3] device buffer
        that the input conversion then reads directly, launched in front of the graph); advances to the next parity / context.
        With ``frames_in_flight`` > 1 the work is enqueued on the context's own stream behind everything enqueued so far on the
        current stream; ``join()`` makes the current stream wait for the frames in flight."""
        k = self.parity
        ext = frame_src is not None
        tag = "ext" if ext else bool(from_frame)
        if self.F == 1:
            if ext:
                self.plans[k].run_input(True, frame_src)
            self._graphed((k, tag), lambda: self._enqueue(k, from_frame, not ext))
            self.parity = 1 - k
            return k
        dev = self.eng.device
        cur, cs = torch.cuda.current_stream(dev), self.cs[k]
        nxt, prev = (k + 1) % self.nctx, (k - 1) % self.nctx
        cs.wait_stream(cur)                                 # the frame was put into plan.frame_in / frame_src on the current stream
        if self._ctx_used[nxt]:
            cs.wait_event(self.ev_frame[nxt])               # context k's previous frame has been matched against (by the frame in context k+1)
        with torch.cuda.stream(cs):
            if ext:
                self.plans[k].run_input(True, frame_src)
            self._graphed((k, tag, 1), lambda: self._enqueue(k, from_frame, not ext, part=1))
            if self._ctx_used[prev]:
                cs.wait_event(self.ev_frame[prev])          # the previous frame's keypoint count / selection / descriptors exist
            self._graphed((k, 2), lambda: self._enqueue(k, part=2))
            self.ev_frame[k].record(cs)
        self._ctx_used[k] = True
        self.parity = nxt
        return k

    def prepare(self, host: bool = True):
        """Capture the CUDA graphs of every frame context (device path, and the host path's variant that reads the staging buffer)
        by running each context once on whatever its input buffers hold; the tracking state is reset afterwards.  Set-up, not work:
        without it the first ``nctx`` frames pay an eager run + capture each."""
        for _ in range(self.nctx):
            self.step_device(True)
        self.join()
        if host:
            # (noise, not a constant image: a flat heat map is the worst case of the keypoint NMS -- every pixel a candidate, one
            # priority chain over the whole frame -- and costs ~30 ms per frame in the single-CTA sweep)
            noise = np.random.RandomState(0).randint(0, 256, (self.B, self.H, self.W, 3)).astype(np.uint8)
            for _ in range(self.nctx):
                self.submit_host(noise)
                self.collect()
        torch.cuda.synchronize(self.eng.device)
        self.reset_tracking()

    def join(self):
        """Make the current stream wait for every frame in flight (no-op with ``frames_in_flight`` = 1: same stream)."""
        if self.F > 1:
            cur = torch.cuda.current_stream(self.eng.device)
            for c in range(self.nctx):
                if self._ctx_used[c]:
                    cur.wait_event(self.ev_frame[c])

    def reset_tracking(self):
        if self.F > 1:
            torch.cuda.synchronize(self.eng.device)
            self._ctx_used = [False] * self.nctx
        for c in self.kcount:
            c.zero_()
        self.parity = 0

    def nms_stats(self) -> np.ndarray:
        """[B,4] int32 of the last frame's box NMS: rows passing objectness, candidates, candidates sorted, path (see the header)."""
        last = (self.parity - 1) % self.nctx
        return self.ws_nms[last][: 16 * self.B].view(torch.int32).view(self.B, 4).cpu().numpy()

    # ---- host boundary -------------------------------------------------------------------------
    def submit_host(self, frames_u8: np.ndarray):
        """Stage the frames in pinned memory and enqueue H2D (copy stream) -> the whole pipeline (current stream) -> D2H of the result
        counts (second copy stream).  At most ``max_outstanding`` (2, or frames_in_flight + 1) submissions may be outstanding:
        ``submit(i+1)`` before ``collect(i)`` lets the host stage / unpack one frame while the GPU works on the next ones, and lets
        the transfers overlap the kernels."""
        if self._n_submit - self._n_collect >= self.nctx:
            raise RuntimeError(f"FramePipeline: {self.nctx} frames already in flight; call collect() first")
        L, dev = _lib.lib(), self.eng.device
        slot = self._n_submit % self.nctx
        h = self._host[slot]
        cur = torch.cuda.current_stream(dev)
        np.copyto(h["frame_np"], np.asarray(frames_u8).reshape(h["frame_np"].shape))
        if h["used"]:
            self.s_in.wait_event(h["ev_read"])              # the staging buffer's previous frame has been consumed
        _lib.check(L.yp_memcpy_async(self.d_frame[slot].data_ptr(), h["frame"].data_ptr(), h["frame"].numel(), C.c_void_p(self.s_in.cuda_stream)))
        h["ev_in"].record(self.s_in)
        cur.wait_event(h["ev_in"])
        k = self.step_device(True, self.d_frame[slot])
        done = cur if self.F == 1 else self.cs[k]
        h["ev_read"].record(done)                           # (conservative: the conversion kernel is the only reader)
        self.s_out.wait_event(h["ev_read"])
        _lib.check(L.yp_memcpy_async(h["counts"].data_ptr(), self.d_counts[k].data_ptr(), 16 * self.B, C.c_void_p(self.s_out.cuda_stream)))
        h["ev_counts"].record(self.s_out)
        h["k"], h["used"] = k, True
        self._n_submit += 1

    def _regrow_and_rerun(self):
        """A frame exceeded an explicitly reduced capacity: grow to the defaults that cannot overflow and redo the frames in
        flight (their host copies are still staged).  The previous frame's keypoints / descriptors are carried over."""
        torch.cuda.synchronize(self.eng.device)
        pending = [(self._host[i % self.nctx]["frame"].numpy().copy(), self._host[i % self.nctx]["k"]) for i in range(self._n_collect, self._n_submit)]
        old = dict(descs_all=self.descs_all, sel=self.sel, kcount=self.kcount)
        self.max_pts, self.nms_cap = self.max_pts_bound(), NMS_CAP
        self._alloc(old)
        self.regrown += 1
        self._n_submit = self._n_collect
        self.parity = pending[0][1]       # (every recorded event has completed: the waits of the re-run frames fall through)
        for frame, _ in pending:
            self.submit_host(frame)

    def collect(self):
        """Wait for the oldest outstanding submit_host and unpack per-image (pts[3,N] f64, desc[D,N] f32, boxes[n,6] f32,
        matches[3,L] f64).  Reads back exactly the rows the counts name."""
        if self._n_collect >= self._n_submit:
            raise RuntimeError("FramePipeline.collect() without a matching submit_host()")
        h = self._host[self._n_collect % self.nctx]
        h["ev_counts"].synchronize()
        cnt = h["counts_np"]
        if cnt[:2].min() < 0:                               # keypoint or box buffer smaller than this frame needs
            self._regrow_and_rerun()
            h = self._host[self._n_collect % self.nctx]
            h["ev_counts"].synchronize()
            cnt = h["counts_np"]
        self._n_collect += 1
        L, k, D = _lib.lib(), h["k"], self.D
        so = C.c_void_p(self.s_res.cuda_stream)       # frame i is complete (its counts arrived): no device-side dependency needed
        nbytes = cnt.nbytes
        for b in range(self.B):
            nk, nb, nm = int(cnt[0, b]), int(cnt[1, b]), int(cnt[2, b])
            for name, src, n, w in (("pts", self.pts[k], nk, 3), ("desc", self.descs[k], nk, D), ("boxes", self.boxes[k], nb, 6), ("matches", self.matches[k], nm, 3)):
                if n > 0:
                    row = src.shape[1] * w * 4 * b
                    _lib.check(L.yp_memcpy_async(h[name].data_ptr() + row, src.data_ptr() + row, n * w * 4, so))
                    nbytes += n * w * 4
        h["ev_done"].record(self.s_res)
        h["ev_done"].synchronize()
        self._d2h_bytes = nbytes
        self.last_threshold_counts = cnt[3].copy()          # pixels >= detection threshold per image (before the NMS)
        out = []
        for b in range(self.B):
            nk, nb, nm = int(cnt[0, b]), int(cnt[1, b]), max(int(cnt[2, b]), 0)
            pts = h["pts_np"][b, :nk].astype(np.float64).T
            desc = h["desc_np"][b, :nk].copy().T            # [D, N] view of a fresh [N, D] copy (same values / shape as the reference's)
            out.append((pts, desc, h["boxes_np"][b, :nb].copy(), h["matches_np"][b, :nm].astype(np.float64).T))
        return out

    def step_host(self, frames_u8: np.ndarray):
        """frames [B,H,W,3] uint8 on the host -> per-image (pts, desc, boxes, matches)."""
        self.submit_host(frames_u8)
        return self.collect()

    def d2h_bytes(self) -> int:
        """Bytes the last collect() read back (counts + exactly the result rows)."""
        return self._d2h_bytes

    def h2d_bytes(self) -> int:
        return self._host[0]["frame"].numel()


class YoloPointFrontend:
    """Drop-in for the reference frontend's per-frame call (src/demo.py:15-230) around an already-built model.

    ``process_img(img, rostpc=None)`` takes a uint8 HxWx3 frame and returns ``(pts [3,N] float64, desc [D,N] float32, [boxes [n,6]])``
    exactly like the reference: crop to multiples of 32, the optional ``crop_resize`` window of the config (crop + ``cv2.resize`` on the
    host, as the reference does it) and the optional per-camera template mask (``self.templates[rostpc]``, src/demo.py:187-192) included.
    """

    def __init__(self, model, config: Optional[dict] = None, filter_pts: bool = True, max_pts: Optional[int] = None, nms_cap: int = NMS_CAP):
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
            if pts.shape[1] >= 2:    # (the reference indexes the first two POINTS here and raises IndexError with fewer)
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
        if int(self.pipeline(H, W).last_threshold_counts[0]) == 0:
            return np.zeros((3, 0)), None, None  # src/demo.py:152-153: no heat-map pixel reaches the detection threshold
        if pts.shape[1] == 0:                    # the border / in-box filters removed every point: src/demo.py:202-203
            pts, obj = self.restore_coords(pts, obj, cth, ctw, resize_fac)
            return pts, np.zeros((self.pipeline(H, W).D, 0)), [obj]
        if self.filter_pts and rostpc:
            pts, desc, dropped = self.template_filter(pts, desc, rostpc)
            if dropped:
                self.last_matches = None          # the device-side matches index the unfiltered sets
        pts, obj = self.restore_coords(pts, obj, cth, ctw, resize_fac)
        return pts, desc, [obj]
