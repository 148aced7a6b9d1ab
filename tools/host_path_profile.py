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
