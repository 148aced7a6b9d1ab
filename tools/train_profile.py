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

# ---- wall-clock phases of one step (synchronised between phases; graphs as in bench.py) ----
import time
from yolopoint_b200 import losses as Lz
from yolopoint_b200.trainer import LAMBDA_DESC, LAMBDA_OBJ
ts2 = TrainStep(m, graph_sample=smp["image"])
for _ in range(3):
    ts2.step(smp)


def phase(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3


acc = {}
for _ in range(3):
    ts2.reducer.zero()
    (semi, desc, obj), t = phase(lambda: ts2._forward(smp["image"], 0)); acc["forward 1"] = acc.get("forward 1", 0) + t
    (semi_w, desc_w, _), t = phase(lambda: ts2._forward(smp["warped_image"], 1)); acc["forward 2"] = acc.get("forward 2", 0) + t
    (lo, _), t = phase(lambda: ts2.obj_loss(obj, smp["box_labels"])); acc["object loss"] = acc.get("object loss", 0) + t
    ld, t = phase(lambda: ts2.det_loss(semi, Lz.labels2Dto3D(smp["labels_2D"]), Lz.getMasks(smp["valid_mask"], ts2.device)) +
                  ts2.det_loss(semi_w, Lz.labels2Dto3D(smp["warped_labels"]), Lz.getMasks(smp["warped_valid_mask"], ts2.device))); acc["detector losses"] = acc.get("detector losses", 0) + t
    lde, t = phase(lambda: Lz.descriptor_loss_sparse(desc, desc_w, smp["warped_valid_mask"], smp["inv_homographies"], **ts2.sparse_cfg)); acc["descriptor loss"] = acc.get("descriptor loss", 0) + t
    loss = ld + LAMBDA_DESC * lde + LAMBDA_OBJ * lo
    _, t = phase(lambda: loss.backward()); acc["backward"] = acc.get("backward", 0) + t
    _, t = phase(lambda: (ts2.reducer.finish(), ts2.opt.step())); acc["all-reduce wait + Adam"] = acc.get("all-reduce wait + Adam", 0) + t
print("wall-clock phases (ms per step, CUDA graphs on, synchronised between phases):")
for k, v in acc.items():
    print(f"  {k:24s} {v / 3:8.2f}")
print(f"  {'sum':24s} {sum(acc.values()) / 3:8.2f}")
