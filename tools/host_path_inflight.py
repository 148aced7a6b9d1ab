"""Host-side cost of FramePipeline.submit_host / collect with several frames in flight (run on the GPU box): wall time per frame of
the two calls, the achieved frames/s, and a cProfile of the loop.  usage: python tools/host_path_inflight.py [frames_in_flight] [tile_policy]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
F = int(sys.argv[1]) if len(sys.argv) > 1 else 6
if len(sys.argv) > 2:
    os.environ["YP_TILE_POLICY"] = sys.argv[2]
sys.argv = [sys.argv[0]]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from yolopoint_b200 import FramePipeline  # noqa: E402
from yolopoint_b200.synth import synthetic_frame  # noqa: E402

model, sd = bench.build_weights("s")
model = model.cuda().eval()
pipe = FramePipeline(model, 1, 640, 640, frames_in_flight=F)
frames = [synthetic_frame(640, 640, s)[None] for s in range(4)]
for i in range(2 * pipe.nctx):
    pipe.step_host(frames[i % 4])


def loop(N, ts=None, tc=None):
    for j in range(F):
        pipe.submit_host(frames[j % 4])
    for i in range(N):
        t0 = time.perf_counter()
        if i + F < N:
            pipe.submit_host(frames[(i + F) % 4])
        t1 = time.perf_counter()
        pipe.collect()
        t2 = time.perf_counter()
        if ts is not None:
            ts.append(t1 - t0); tc.append(t2 - t1)


ts, tc = [], []
torch.cuda.synchronize()
t00 = time.perf_counter()
loop(300, ts, tc)
torch.cuda.synchronize()
tot = time.perf_counter() - t00
print(f"frames_in_flight {F}: e2e {300 / tot:.1f} frames/s | submit_host {1e6 * np.median(ts[:-F]):.0f} us, collect {1e6 * np.median(tc):.0f} us (median wall per frame)")
pr = cProfile.Profile()
pr.enable()
loop(200)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
