"""Per-layer time table of the conv launch list (run on the GPU box).

  python tools/layer_times.py [--version s] [--size 640 640] [--batch 1] [--precision fp32] [--no-tuning]

Every conv launch of the ShapePlan is timed alone: 20 back-to-back launches captured in a CUDA graph, CUDA events on the
launching stream, best of 3 (operands therefore come from L2 when they fit).  Prints us, CTAs, GFLOP and TFLOP/s per layer
and the total; the sum over-estimates the in-network time where head lanes overlap the trunk."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

from yolopoint_b200 import Model, _lib  # noqa: E402
from yolopoint_b200.synth import perturb_state_dict  # noqa: E402
from tune_conv import time_desc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--version", default="s")
    ap.add_argument("--size", type=int, nargs=2, default=[640, 640])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--no-tuning", action="store_true")
    args = ap.parse_args()
    H, W = args.size
    L = _lib.lib(require_device=True)
    torch.manual_seed(0)
    m = Model(names=[str(i) for i in range(80)], version=args.version, precision=args.precision)
    m.load_state_dict(perturb_state_dict(m.state_dict(), 0, args.version))
    m = m.cuda().eval()
    eng = m.engine()
    eng.use_tuning = not args.no_tuning
    plan = eng.plan(args.batch, H, W)
    tot_us = tot_gf = 0.0
    rows = []
    for op, d in plan.conv_descs:
        t = time_desc(L, d)
        Ho, Wo = d.in_.H // op.s, d.in_.W // op.s
        gf = 2.0 * args.batch * Ho * Wo * op.cout * op.src.C * op.k * op.k / 1e9
        tot_us += t
        tot_gf += gf
        rows.append((t, "+".join(op.names), op, d, gf))
        print(f"{'+'.join(op.names):44s} k{op.k}s{op.s} {op.src.C:4d}->{op.cout:4d} @{d.in_.H}x{d.in_.W} tile_n={d.tile_n} split_k={d.split_k} "
              f"{t:8.2f} us {gf:8.3f} GF {gf / t * 1e3:8.1f} TF/s", flush=True)
    print(f"total {tot_us:.1f} us, {tot_gf:.2f} GFLOP (padded channels), {tot_gf / tot_us * 1e3:.1f} TF/s")


if __name__ == "__main__":
    main()
