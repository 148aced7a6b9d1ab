"""ncu target: a few whole-frame steps (eager launches, no graph) so that -k regex:<kernel> can pick the post-processing kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

sys.argv = [sys.argv[0]]
import bench  # noqa: E402
from yolopoint_b200 import FramePipeline  # noqa: E402
from yolopoint_b200.synth import synthetic_frame  # noqa: E402

model, sd = bench.build_weights("s")
model = model.cuda().eval()
model.engine().use_graphs = False
pipe = FramePipeline(model, 1, 640, 640)
for s in range(4):
    pipe.step_host(synthetic_frame(640, 640, s % 4)[None])
torch.cuda.synchronize()
