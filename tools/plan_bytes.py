"""Algorithmic (compulsory) bytes of the conv launches of one pass, from the launch plan alone (no GPU): every launch reads its
input slice (+ residual) and its packed weights once and writes its outputs once, in the storage format of the buffers
((hi, lo) fp32 planes = 8 B per element on the 3xTF32 path, bf16 = 2 B, plain fp32 network outputs = 4 B).

    python tools/plan_bytes.py [version] [B] [H] [W] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolopoint_b200._lib import YP_FMT_BF16, YP_FMT_F32, YP_FMT_F32X2  # noqa: E402
from yolopoint_b200.engine import NetPlan  # noqa: E402


def plan_bytes(version="s", B=1, H=640, W=640, precision="fp32", model_name="YOLOPoint"):
    net = NetPlan(version, 80, precision, model_name)
    es = {YP_FMT_BF16: 2, YP_FMT_F32X2: 8, YP_FMT_F32: 4}
    wes = 8 if precision == "fp32" else 2
    rows = []
    for op in net.conv_ops():
        def nbytes(ref):
            # a 2x-upsampled destination view addresses the low-resolution grid and its store writes 4 pixels per element:
            # the bytes are those of the slice at the BUFFER's resolution either way
            lvl, _, fmt = net.bufs[ref.buf]
            return B * (H >> lvl) * (W >> lvl) * ref.C * es[fmt]
        rd = nbytes(op.src) + (nbytes(op.residual) if op.residual is not None else 0)
        wr = sum(nbytes(d) for d in op.dst)
        wt = op.cout * op.src.C * op.k * op.k * wes
        rows.append(("+".join(op.names), rd, wt, wr))
    return rows


if __name__ == "__main__":
    a = sys.argv[1:]
    rows = plan_bytes(a[0] if a else "s", *(int(v) for v in a[1:4]), *(a[4:5])) if len(a) > 1 else plan_bytes(*a)
    rd, wt, wr = (sum(r[i] for r in rows) for i in (1, 2, 3))
    print(f"{len(rows)} conv launches: activations read {rd / 1e6:.1f} MB, weights read {wt / 1e6:.1f} MB, written {wr / 1e6:.1f} MB; "
          f"per launch {(rd + wt) / len(rows) / 1e6:.2f} MB read + {wr / len(rows) / 1e6:.2f} MB written")
