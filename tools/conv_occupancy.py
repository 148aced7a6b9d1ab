"""Whole-grid picture of one tcgen05 conv launch (debug aid, run on the GPU box): per-CTA (SM, start, end) from %globaltimer,
plus the pipeline stamps of CTA 0.  Prints the launch span, CTA lifetime statistics, CTAs per SM and the idle gaps."""
import ctypes as C
import os
import sys

os.environ.setdefault("YP_CONV_PERSIST", "0")   # the recorder lives in the one-tile-per-CTA kernel; the persistent variant is timed by tools/train_conv_times.py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from yolopoint_b200 import _lib  # noqa: E402
from yolopoint_b200._lib import YP_ALGO_TCGEN05, YP_FMT_BF16, YP_FMT_F32X2  # noqa: E402
import test_gpu_conv as T  # noqa: E402

SHAPES = [
    dict(B=8, H=80, W=80, Cin=128, Cout=128, k=3, s=1, act=False, nobias=True),     # YOLOPoint-L training layer, 15x per pass
    dict(B=8, H=80, W=80, Cin=128, Cout=128, k=1, s=1, act=False, nobias=True),
    dict(B=8, H=160, W=160, Cin=64, Cout=64, k=3, s=1, act=False, nobias=True),
    dict(B=8, H=160, W=160, Cin=64, Cout=128, k=3, s=2, act=False, nobias=True),
    dict(B=1, H=80, W=80, Cin=64, Cout=64, k=3, s=1, res=True),                     # YOLOPoint-S batch 1
    dict(B=1, H=80, W=80, Cin=128, Cout=128, k=1, s=1),
    dict(B=1, H=160, W=160, Cin=32, Cout=32, k=1, s=1),                             # smallest layer of YOLOPoint-S
]


def main():
    L = _lib.lib(require_device=True)
    buf = torch.zeros(512 + 3 * 20000, dtype=torch.int64, device="cuda")
    fmts = [YP_FMT_BF16, YP_FMT_F32X2] if len(sys.argv) < 2 else [dict(bf16=YP_FMT_BF16, fp32=YP_FMT_F32X2)[sys.argv[1]]]
    for fmt in fmts:
        for c in SHAPES:
            for rep in range(3):
                buf.zero_()
                L.yp_debug_conv_timeline(C.c_void_p(buf.data_ptr()))
                try:
                    T.run_case(c, fmt, YP_ALGO_TCGEN05)
                except AssertionError as e:
                    print("numerics:", e)
                L.yp_debug_conv_timeline(None)
            t = buf.cpu().numpy()
            rec = t[512:].reshape(-1, 3)
            rec = rec[rec[:, 1] > 0]
            if len(rec) == 0:
                print(f"\n== fmt={'bf16' if fmt == YP_FMT_BF16 else 'f32x2'} {c}: no per-CTA records (kernel without recorder)")
                continue
            sm, st, en = rec[:, 0], rec[:, 1], rec[:, 2]
            t0 = st.min()
            life = (en - st) / 1e3
            span = (en.max() - t0) / 1e3
            per_sm = np.bincount(sm.astype(int), minlength=148)
            print(f"\n== fmt={'bf16' if fmt == YP_FMT_BF16 else 'f32x2'} {c}")
            print(f"  CTAs {len(rec)}  launch span {span:.1f} us  CTA lifetime us: min {life.min():.1f} median {np.median(life):.1f} max {life.max():.1f}")
            print(f"  CTAs per SM: min {per_sm.min()} max {per_sm.max()}  SMs used {(per_sm > 0).sum()}  last CTA start at {(st.max() - t0) / 1e3:.1f} us")
            # how many CTAs are resident over time (10 samples)
            pts = np.linspace(0, span, 11)[:-1] + span / 20
            res = [int(((st - t0) / 1e3 <= p).sum() - ((en - t0) / 1e3 <= p).sum()) for p in pts]
            print("  resident CTAs over the span:", res)
            mhz = 1965.0
            s0 = t[0]
            us = lambda v: (v - s0) / mhz if v else float("nan")
            nkb = sum(1 for v in t[8:104] if v)
            print(f"  CTA0: prologue {us(t[1]):.2f} | accum ready {us(t[2]):.2f} | stores drained {us(t[3]):.2f} | dealloc {us(t[5]):.2f}  (K-loop units {nkb})")
            prod = [us(v) for v in t[8:8 + nkb]]
            full = [us(v) for v in t[104:104 + nkb]]
            iss = [us(v) for v in t[200:200 + nkb]]
            if t[320]:
                print(f"    epilogue c0: first barrier {us(t[320]):.2f}  residual loaded {us(t[321]):.2f}  accumulators loaded {us(t[322]):.2f}")
            for ci in range(4):
                aa, bb, dd = t[300 + 4 * ci], t[301 + 4 * ci], t[302 + 4 * ci]
                if aa:
                    print(f"    chunk {ci}: math done {us(aa):7.2f}  staged {us(bb):7.2f}  store issued {us(dd):7.2f}")
            for i in list(range(min(nkb, 4))) + ([nkb - 1] if nkb > 4 else []):
                print(f"    unit {i:3d}: TMA issued {prod[i]:7.2f}  landed {full[i]:7.2f}  MMAs issued {iss[i]:7.2f}")


if __name__ == "__main__":
    main()
