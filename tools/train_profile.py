"""Kernel-time breakdown of one training step (torch.profiler, eager launches): python tools/train_profile.py [--version l] [--batch 8]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from yolopoint_b200 import Model  # noqa: E402
from yolopoint_b200.trainer import TrainStep, synthetic_sample  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--version", default="l")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, nargs=2, default=[640, 640])
ap.add_argument("--backend", default="b200")
args = ap.parse_args()
torch.manual_seed(0)
m = Model(names=[str(i) for i in range(80)], version=args.version).cuda().train()
m.train_backend = args.backend
ts = TrainStep(m)
smp = {k: v.cuda() for k, v in synthetic_sample(args.batch, args.size[0], args.size[1], 0).items()}
for _ in range(3):
    ts.step(smp)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        ts.step(smp)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time per step {tot / 2e3:.2f} ms")
for e in rows[:40]:
    print(f"{e.device_time_total / 2e3:9.3f} ms {100 * e.device_time_total / tot:5.1f}%  n={e.count // 2:5d}  {e.key[:110]}")
