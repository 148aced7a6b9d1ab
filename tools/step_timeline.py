"""Timeline of ONE whole-frame step inside its CUDA graph (run on the GPU box): every kernel with start / duration / stream from the
CUPTI activity records of torch.profiler, relative to the first kernel of the step.  Unlike the ncu launch list (serialised, cold
caches) this shows the step as it runs: lanes overlapping, gaps between dependent kernels, what is on the critical path.

  python tools/step_timeline.py [--workload s640] [--precision fp32] [--out gpurun_out/timeline.txt]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="s640")
ap.add_argument("--precision", default="fp32")
ap.add_argument("--out", default="")
ap.add_argument("--net-only", action="store_true")
args = ap.parse_args()
sys.argv = [sys.argv[0]]
import bench  # noqa: E402
from yolopoint_b200 import FramePipeline  # noqa: E402
from yolopoint_b200.synth import synthetic_frame  # noqa: E402

version, H, W, per_gpu = bench.WORKLOADS[args.workload]
model, sd = bench.build_weights(version, bench.MODEL_NAME.get(args.workload, "YOLOPoint"))
model.precision = args.precision
model = model.cuda().eval()
pipe = FramePipeline(model, per_gpu, H, W)
frames = [torch.from_numpy(synthetic_frame(H, W, s)).cuda() for s in range(4)]


def step(i):
    for b in range(per_gpu):
        pipe.plan.frame_in[b].copy_(frames[(i + b) % 4])
    if args.net_only:
        pipe.plan.graphed("net_only", pipe.plan.run_net)
    else:
        pipe.step_device(True)


for i in range(6):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
import json  # noqa: E402
import re  # noqa: E402
import tempfile  # noqa: E402

tmp = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(tmp)
tr = json.load(open(tmp))["traceEvents"]
ev = [e for e in tr if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and e.get("dur", 0) > 0]
ev.sort(key=lambda e: e["ts"])
# split into steps at the input copies (Memcpy DtoD of the frame) and keep the last one
starts = [i for i, e in enumerate(ev) if e.get("cat") == "gpu_memcpy"]
first = starts[-per_gpu] if len(starts) >= per_gpu else 0
ev = ev[first:]
t0 = ev[0]["ts"]
streams = {}
lines = []
end_max = 0.0


def short(name):
    name = name.replace("void ", "").replace("yp::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*", "", name)
    return name[:70]


for e in ev:
    s = streams.setdefault(e.get("args", {}).get("stream"), len(streams))
    st, du = e["ts"] - t0, e["dur"]
    end_max = max(end_max, st + du)
    g = e.get("args", {}).get("grid")
    lines.append(f"{st:9.1f} {du:7.1f}  s{s}  {short(e['name']):70s} grid {g}")
hdr = f"# one step of {args.workload} ({args.precision}){' network only' if args.net_only else ''}: {len(ev)} GPU activities, span {end_max:.1f} us\n#  start us   dur us  stream  kernel"
txt = hdr + "\n" + "\n".join(lines)
print(txt)
if args.out:
    open(args.out, "w").write(txt + "\n")
