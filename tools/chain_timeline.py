"""Per-operation timeline of the layer chains (YP_CHAIN_DEBUG recorder of conv_chain_kernel) on a B200.

For every operation of every segment: when its first item started, when the dependencies of its last-starting item were met, and
the phases of the slowest CTA (prologue done, accumulators complete, stores complete, completion published), all in microseconds
relative to the first stamp of the segment.

usage: YP_CHAIN_DEBUG=1 python tools/chain_timeline.py [version] [B] [H] [W]
"""
import ctypes as C
import os
import sys

os.environ.setdefault("YP_CHAIN_DEBUG", "1")
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolopoint_b200 import Model, _lib  # noqa: E402
from yolopoint_b200.engine import Engine  # noqa: E402
from yolopoint_b200.synth import perturb_state_dict  # noqa: E402


def main():
    ver = sys.argv[1] if len(sys.argv) > 1 else "s"
    B, H, W = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (1, 640, 640)
    torch.manual_seed(0)
    m = Model(names=[str(i) for i in range(80)], version=ver)
    sd = perturb_state_dict(m.state_dict(), 0, ver)
    dev = torch.device("cuda:0")
    eng = Engine(sd, ver, 80, dev, chain=True)
    p = eng.plan(B, H, W)
    x = torch.rand(B, 3, H, W, device=dev)
    for _ in range(3):
        eng.forward(x)
    torch.cuda.synchronize()
    L = _lib.lib()
    for si, sg in enumerate(p.chain):
        # the segment alone, a few times, so that the stamps are those of a warm run
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            _lib.check(L.yp_conv_chain_launch(sg["handle"], st))
        torch.cuda.synchronize()
        ptr, n_ops, n_ctas = C.c_void_p(), C.c_int32(), C.c_int32()
        _lib.check(L.yp_debug_conv_chain_timeline(sg["handle"], 0, C.byref(ptr), C.byref(n_ops), C.byref(n_ctas)))
        n = n_ops.value * n_ctas.value * 16
        host = (C.c_int64 * n)()
        torch.cuda.synchronize()
        _lib.check(L.yp_memcpy_async(host, ptr, n * 8, st))
        torch.cuda.synchronize()
        t = torch.tensor(list(host), dtype=torch.int64).view(n_ops.value, n_ctas.value, 16).double()
        t0 = t[t > 0].min()
        print(f"== segment {si}: {n_ops.value} ops, {sg['items']} items, span {(t.max() - t0) / 1e3:.1f} us")
        print("  op                                   ctas | first start | deps met (last) | slowest CTA: start, deps, prologue, tma0 issued, landed, kb0 issued, all issued, accum, stores, end, published | layer span")
        for l in range(n_ops.value):
            r = t[l]
            used = r[:, 0] > 0
            if not used.any():
                continue
            r = (r[used] - t0) / 1e3
            slow = int(r[:, 7].argmax())
            name = sg["names"][l][:36]
            base = r[:, 6].max()
            ph = " ".join(f"{float(r[slow, k] - base):6.2f}" for k in (0, 6, 1, 8, 9, 10, 11, 2, 3, 4, 7))
            print(f"  {name:36s} {int(used.sum()):4d} | {r[:, 0].min():8.2f}    | {base:8.2f}        | {ph} | {r[:, 7].max() - base:6.2f}")


if __name__ == "__main__":
    main()
