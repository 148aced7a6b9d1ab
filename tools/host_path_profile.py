"""Host-side cost of the public per-frame API (run on the GPU box): wall time of FramePipeline.submit_host / collect per frame with the
host running one frame ahead, next to the device time of one step.  Tells whether `e2e` is bound by the GPU or by the Python host."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

sys.argv = [sys.argv[0]]
import bench  # noqa: E402
from yolopoint_b200 import FramePipeline  # noqa: E402
from yolopoint_b200.synth import synthetic_frame  # noqa: E402

model, sd = bench.build_weights("s")
model = model.cuda().eval()
pipe = FramePipeline(model, 1, 640, 640)
frames = [synthetic_frame(640, 640, s)[None] for s in range(4)]
for i in range(4):
    pipe.step_host(frames[i % 4])
N = 200
ts, tc = [], []
torch.cuda.synchronize()
t00 = time.perf_counter()
pipe.submit_host(frames[0])
for i in range(N):
    t0 = time.perf_counter()
    if i + 1 < N:
        pipe.submit_host(frames[(i + 1) % 4])
    t1 = time.perf_counter()
    pipe.collect()
    t2 = time.perf_counter()
    ts.append(t1 - t0); tc.append(t2 - t1)
torch.cuda.synchronize()
tot = time.perf_counter() - t00
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(50):
    pipe.step_device(True)
e1.record(); torch.cuda.synchronize()
print(f"e2e {N / tot:.1f} frames/s | submit_host {1e6 * np.median(ts):.0f} us, collect {1e6 * np.median(tc):.0f} us (median wall time per frame) | "
      f"device step {e0.elapsed_time(e1) / 50 * 1e3:.0f} us | d2h bytes {pipe.d2h_bytes()}")
# device-side picture of the same loop: per frame, time from the first kernel to the last (busy) and from one frame's end to the
# next frame's start (gap), from timing events recorded on the compute stream around step_device
orig = pipe.step_device
marks = []


def timed_step(*a, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k = orig(*a, **kw)
    e1.record()
    marks.append((e0, e1))
    return k


pipe.step_device = timed_step
pipe.submit_host(frames[0])
for i in range(60):
    pipe.submit_host(frames[(i + 1) % 4])
    pipe.collect()
pipe.collect()
torch.cuda.synchronize()
busy = [a.elapsed_time(b) * 1e3 for a, b in marks[5:]]
gap = [marks[i][1].elapsed_time(marks[i + 1][0]) * 1e3 for i in range(5, len(marks) - 1)]
print(f"device: busy {np.median(busy):.0f} us per frame (min {min(busy):.0f}, max {max(busy):.0f}), gap between frames {np.median(gap):.0f} us (max {max(gap):.0f})")
# the same picture for the device-resident loop (bench.py's `value`): frame already in HBM, host far ahead
marks.clear()
pool = [torch.from_numpy(f).cuda() for f in frames]
for i in range(60):
    pipe.plan.frame_in.copy_(pool[i % 4])
    timed_step(True)
torch.cuda.synchronize()
busy = [a.elapsed_time(b) * 1e3 for a, b in marks[5:]]
gap = [marks[i][1].elapsed_time(marks[i + 1][0]) * 1e3 for i in range(5, len(marks) - 1)]
print(f"device loop: busy {np.median(busy):.0f} us per frame, gap {np.median(gap):.0f} us (max {max(gap):.0f})")
# and for the host loop with the host TWO frames ahead is impossible (two staging slots); instead: host loop without reading results
marks.clear()
t0 = time.perf_counter()
for i in range(60):
    h = pipe._host[i % 2]
    if pipe._n_submit - pipe._n_collect >= 2:
        pipe._host[pipe._n_collect % 2]["ev_counts"].synchronize(); pipe._n_collect += 1
    pipe.step_device = timed_step
    pipe.submit_host(frames[i % 4])
torch.cuda.synchronize()
pipe._n_collect = pipe._n_submit
busy = [a.elapsed_time(b) * 1e3 for a, b in marks[5:]]
gap = [marks[i][1].elapsed_time(marks[i + 1][0]) * 1e3 for i in range(5, len(marks) - 1)]
print(f"host loop, counts only: busy {np.median(busy):.0f} us per frame, gap {np.median(gap):.0f} us (max {max(gap):.0f}), {60 / (time.perf_counter() - t0):.0f} frames/s")
pipe.step_device = orig
import cProfile, pstats  # noqa: E402
pr = cProfile.Profile()
pr.enable()
pipe.submit_host(frames[0])
for i in range(100):
    pipe.submit_host(frames[(i + 1) % 4])
    pipe.collect()
pipe.collect()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
