"""How far the engine's fp32 (3xTF32) network outputs are from a float64 evaluation, next to the fp32 reference arithmetic (run on the
GPU box).  Usage: python tools/precision_probe.py [--version s] [--size 640 640] [--no-tuning]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import yolopoint_oracle as O  # noqa: E402
from yolopoint_b200 import Model  # noqa: E402
from yolopoint_b200.synth import perturb_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--version", default="s")
ap.add_argument("--size", type=int, nargs=2, default=[640, 640])
ap.add_argument("--no-tuning", action="store_true")
args = ap.parse_args()
H, W = args.size
torch.manual_seed(0)
m = Model(names=[str(i) for i in range(80)], version=args.version)
sd = perturb_state_dict(m.state_dict(), 0, args.version)
m.load_state_dict(sd)
m = m.cuda().eval()
m.engine().use_tuning = not args.no_tuning
x = torch.from_numpy(np.random.RandomState(11).rand(1, 3, H, W).astype(np.float32))
out = m(x.cuda())
r32 = O.OracleNet(sd, args.version, 80).forward(x)
r64 = O.OracleNet(sd, args.version, 80, dtype=torch.float64).forward(x)
tag = f"{args.version} {H}x{W} tuning={'off' if args.no_tuning else 'on'} env={ {k: v for k, v in os.environ.items() if k.startswith('YP_')} }"
print(tag)
for name, g, a, e in [("semi", out["semi"], r32["semi"], r64["semi"]), ("desc", out["desc"], r32["desc"], r64["desc"]),
                      ("pred", out["objects"][0], r32["objects"][0], r64["objects"][0])] + \
                     [(f"raw{i}", out["objects"][1][i], r32["objects"][1][i], r64["objects"][1][i]) for i in range(3)]:
    dg, dr = (g.double().cpu() - e).abs(), (a.double() - e).abs()
    bias = float(((g.double().cpu() - e) * torch.sign(e)).mean())
    print(f"  {name:5s} |exact|max {float(e.abs().max()):9.3f}  engine: max {float(dg.max()):.3e} mean {float(dg.mean()):.3e} signed-toward-zero mean {bias:+.3e} | "
          f"fp32 reference: max {float(dr.max()):.3e} mean {float(dr.mean()):.3e}")
