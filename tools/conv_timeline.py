"""Per-event timeline of one CTA of the tcgen05 conv kernel (debug aid, run on the GPU box).
Prints clock64 deltas (in SM cycles and microseconds at the reported clock) for several layer shapes."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from yolopoint_b200 import _lib  # noqa: E402
from yolopoint_b200._lib import YP_ALGO_TCGEN05, YP_FMT_BF16, YP_FMT_F32X2  # noqa: E402
import test_gpu_conv as T  # noqa: E402

SHAPES = [
    dict(B=1, H=80, W=80, Cin=64, Cout=64, k=1, s=1),
    dict(B=1, H=80, W=80, Cin=64, Cout=64, k=3, s=1, res=True),
    dict(B=1, H=20, W=20, Cin=256, Cout=256, k=3, s=1, res=True),
    dict(B=1, H=20, W=20, Cin=1024, Cout=512, k=1, s=1),
    dict(B=1, H=320, W=320, Cin=16, Cout=32, k=3, s=1),
    dict(B=4, H=23, W=40, Cin=384, Cout=384, k=3, s=1, res=True),      # YOLOPoint-M stride-32 layer, batch 4 (192 CTAs)
    dict(B=4, H=46, W=80, Cin=192, Cout=192, k=3, s=1, res=True),
]


def main():
    L = _lib.lib(require_device=True)
    buf = torch.zeros(512, dtype=torch.int64, device="cuda")
    mhz = 1965.0
    for fmt in (YP_FMT_F32X2, YP_FMT_BF16):
        for c in SHAPES:
            for rep in range(3):   # third repetition is warm
                buf.zero_()
                L.yp_debug_conv_timeline(C.c_void_p(buf.data_ptr()))
                try:
                    T.run_case(c, fmt, YP_ALGO_TCGEN05)
                except AssertionError as e:
                    print("numerics:", e)
                L.yp_debug_conv_timeline(None)
            t = buf.cpu().tolist()
            t0 = t[0]
            us = lambda v: (v - t0) / mhz if v else float("nan")
            nkb = sum(1 for v in t[8:104] if v)
            print(f"\n== fmt={'f32x2' if fmt == 0 else 'bf16'} {c}  k-blocks={nkb}")
            print(f"  prologue done {us(t[1]):7.2f} us | accum ready {us(t[2]):7.2f} | stores drained {us(t[3]):7.2f} | final sync {us(t[4]):7.2f} | dealloc {us(t[5]):7.2f}")
            prod = [us(v) for v in t[8:8 + nkb]]
            full = [us(v) for v in t[104:104 + nkb]]
            iss = [us(v) for v in t[200:200 + nkb]]
            show = list(range(min(nkb, 6))) + ([nkb - 2, nkb - 1] if nkb > 8 else [])
            for i in show:
                print(f"  kb {i:3d}: TMA issued {prod[i]:7.2f}  data landed(MMA saw full) {full[i]:7.2f}  MMAs issued {iss[i]:7.2f}")
            for ci in range(8):
                a, b, d = t[300 + 4 * ci], t[301 + 4 * ci], t[302 + 4 * ci]
                if a:
                    print(f"  chunk {ci}: math done {us(a):7.2f}  staged {us(b):7.2f}  store issued+prev drained {us(d):7.2f}")


if __name__ == "__main__":
    main()
