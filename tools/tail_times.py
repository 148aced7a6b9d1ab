"""In-graph time of the two post-processing stages whose inputs are network outputs (run on the GPU box): the fused Detect decode + box
NMS (on the critical path after the last conv) and heatmap + keypoint NMS rounds (keypoint lane, under the detection branch), each
captured as a CUDA graph and replayed alone (CUDA events, best of 5 x 20 replays).  The later stages (collect / emit, sampling, match)
consume state the earlier ones produce and are only meaningful inside the whole pipeline (bench.py: step time minus network time)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

sys.argv = [sys.argv[0]]
import bench  # noqa: E402
from yolopoint_b200 import FramePipeline, _lib  # noqa: E402
from yolopoint_b200.synth import synthetic_frame  # noqa: E402

model, sd = bench.build_weights("s")
model = model.cuda().eval()
pipe = FramePipeline(model, 1, 640, 640)
for s in range(3):
    pipe.step_host(synthetic_frame(640, 640, s)[None])
L, p = _lib.lib(), pipe.plan
B, H, W, cfg, k = 1, 640, 640, pipe.cfg, 0


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def stage_nms():
    dets = [p.bufs[f"det{i}"] for i in range(3)]
    lg = (C.c_void_p * 3)(*[d.data_ptr() for d in dets])
    ny = (C.c_int32 * 3)(*[d.shape[2] for d in dets]); nx = (C.c_int32 * 3)(*[d.shape[3] for d in dets])
    ldc = (C.c_int32 * 3)(*[d.shape[4] for d in dets])
    strd = (C.c_float * 3)(*[float(v) for v in pipe.eng.stride])
    anc = (C.c_float * 18)(*[float(v) for row in pipe.eng.anchors_px for v in row])
    _lib.check(L.yp_detect_nms(lg, ny, nx, ldc, strd, anc, B, 3, pipe.eng.net.no, C.byref(pipe.nms_params), pipe.nms_cap, pipe.boxes[0].data_ptr(),
                               pipe.bcount[0].data_ptr(), pipe.ws_nms.data_ptr(), pipe.ws_nms.numel(), 0, st()))


def stage_heat_kpnms():
    semi = p.bufs["semi"][0]
    sB, sH, sW, sC = semi.stride()
    _lib.check(L.yp_heatmap(semi.data_ptr(), B, H // 8, W // 8, sB, sC, sH, sW, pipe.heat_variant, pipe.heat.data_ptr(), st()))
    _lib.check(L.yp_keypoints_nms(pipe.heat.data_ptr(), B, H, W, float(cfg["detection_threshold"]), int(cfg["nms"]), pipe.max_pts, pipe.ws_kp.data_ptr(),
                                  pipe.ws_kp.numel(), st()))


def time_graph(fn, reps=20):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay(); s.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); g.replay(); e1.record(s); s.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


def stage_nms_prescanned():
    dets = [p.bufs[f"det{i}"] for i in range(3)]
    lg = (C.c_void_p * 3)(*[d.data_ptr() for d in dets])
    ny = (C.c_int32 * 3)(*[d.shape[2] for d in dets]); nx = (C.c_int32 * 3)(*[d.shape[3] for d in dets])
    ldc = (C.c_int32 * 3)(*[d.shape[4] for d in dets])
    strd = (C.c_float * 3)(*[float(v) for v in pipe.eng.stride])
    anc = (C.c_float * 18)(*[float(v) for row in pipe.eng.anchors_px for v in row])
    for lvl in (0, 1):
        _lib.check(L.yp_detect_prescan(lg, ny, nx, ldc, strd, anc, B, 3, pipe.eng.net.no, C.byref(pipe.nms_params), pipe.nms_cap, lvl,
                                       pipe.ws_nms.data_ptr(), pipe.ws_nms.numel(), st()))
    _lib.check(L.yp_detect_nms(lg, ny, nx, ldc, strd, anc, B, 3, pipe.eng.net.no, C.byref(pipe.nms_params), pipe.nms_cap, pipe.boxes[0].data_ptr(),
                               pipe.bcount[0].data_ptr(), pipe.ws_nms.data_ptr(), pipe.ws_nms.numel(), 3, st()))


for name, fn in (("Detect decode + box NMS (1 kernel, scans all levels)", stage_nms), ("prescan levels 0, 1 (2 kernels) + box NMS", stage_nms_prescanned),
                 ("heatmap + keypoint NMS rounds (side lane)", stage_heat_kpnms)):
    print(f"{name:55s} {time_graph(fn):8.2f} us")
