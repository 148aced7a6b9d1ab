"""Layer-by-layer check of the conv launch list of a ShapePlan against torch (run on the GPU box).

  python tools/layer_check.py [--version s] [--size 640 640] [--batch 1] [--precision fp32] [--no-tuning]

Runs the launches one at a time (eagerly, single stream); before each conv launch the source / residual slices are
snapshotted, after it the destination slice is compared with F.conv2d (fp64) on the snapshot.  Prints one line per
layer with the launch geometry the planner chose, flags layers whose relative error exceeds the tolerance."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from yolopoint_b200 import Model, _lib  # noqa: E402
from yolopoint_b200.engine import ConvOp  # noqa: E402
from yolopoint_b200.synth import perturb_state_dict  # noqa: E402


def read_slice(plan, ref):
    t = plan.bufs[ref.buf]                       # [planes, B, H, W, C]
    x = t[..., ref.c_off:ref.c_off + ref.C].double().sum(0)
    if ref.upsample == 2:
        x = x[:, ::2, ::2]
    return x                                     # [B, H, W, C]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--version", default="s")
    ap.add_argument("--size", type=int, nargs=2, default=[640, 640])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--no-tuning", action="store_true")
    ap.add_argument("--tol", type=float, default=None)
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
    tol = args.tol if args.tol is not None else (2e-5 if args.precision == "fp32" else 2e-2)
    plan.frame_in.copy_(torch.randint(0, 256, plan.frame_in.shape, dtype=torch.uint8, device="cuda"))
    plan.run_input(True)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    convs = iter(plan.conv_descs)
    bad = 0
    for op, (lane, f, _name) in zip(eng.net.ops, plan.launches):
        if not isinstance(op, ConvOp):
            f(st)
            continue
        _, d = next(convs)
        x = read_slice(plan, op.src).permute(0, 3, 1, 2)
        res = read_slice(plan, op.residual) if op.residual is not None else None
        w, b = eng.weights[op.names]
        wd = w.double().sum(0).view(op.cout, op.k * op.k, op.src.C).permute(0, 2, 1).reshape(op.cout, op.src.C, op.k, op.k)
        f(st)
        torch.cuda.synchronize()
        y = F.conv2d(x, wd, None if b is None else b.double(), stride=op.s, padding=op.k // 2)
        if op.act:
            y = y * torch.sigmoid(y)
        y = y.permute(0, 2, 3, 1)
        if res is not None:
            y = y + res
        if op.l2norm:
            y = y / y.norm(dim=-1, keepdim=True)
        errs = []
        for ds in op.dst:
            got = read_slice(plan, ds)
            errs.append(float((got - y).abs().max() / y.abs().max().clamp_min(1e-30)))
        e = max(errs)
        flag = "  <-- BAD" if not (e < tol) else ""
        bad += bool(flag)
        print(f"{'+'.join(op.names):44s} k{op.k}s{op.s} {op.src.C:4d}->{op.cout:4d} @{d.in_.H}x{d.in_.W} tile_n={d.tile_n} split_k={d.split_k} "
              f"res={int(res is not None)} err={e:.2e}{flag}", flush=True)
    print("bad layers:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
