"""Per-shape times of the training-step convolutions (forward, data gradient, weight gradient) on the GPU box.

  python tools/train_conv_times.py [--version l] [--size 640 640] [--batch 8]

The distinct conv shapes are collected from one train-mode forward of the module tree; each operation is then timed alone: 10 calls captured in
a CUDA graph and replayed (CUDA events on the launching stream, best of 3) next to cuDNN's bf16 kernels for the same operation."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from yolopoint_b200 import Model, train as T  # noqa: E402


def timeit(fn, reps=10):
    """GPU time per call: `reps` calls captured in a CUDA graph (no Python / ctypes / tensor-map encoding time between the
    launches, as in the graph-replayed training step), replayed 3 times, best."""
    fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn()
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
        g.replay()
        st.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            g.replay()
            e1.record(st)
            st.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--version", default="l")
    ap.add_argument("--size", type=int, nargs=2, default=[640, 640])
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()
    H, W = args.size
    torch.manual_seed(0)
    m = Model(names=[str(i) for i in range(80)], version=args.version).cuda().train()
    m.train_backend = "cudnn_bf16"
    shapes = {}

    def hook(mod, inp, out):
        x = inp[0]
        key = (x.shape[1], mod.out_channels, mod.kernel_size[0], mod.stride[0], x.shape[2], x.shape[3])
        shapes[key] = shapes.get(key, 0) + 1
    hs = [mod.register_forward_hook(hook) for mod in m.modules() if isinstance(mod, torch.nn.Conv2d)]
    with torch.no_grad():
        m(torch.rand(1, 3, H, W, device="cuda"))
    for h in hs:
        h.remove()
    B = args.batch
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0, "cudnn_fwd": 0.0, "cudnn_dgrad": 0.0, "cudnn_wgrad": 0.0}
    tot_gf = 0.0
    print(f"{'Ci':>5} {'Co':>5} k s {'HxW':>9} cnt | {'GF':>7} | fwd us (TF/s) | dgrad us (TF/s) | wgrad us (TF/s) | cuDNN fwd / dgrad / wgrad us")
    for (ci, co, k, s, h, w), cnt in sorted(shapes.items(), key=lambda kv: -kv[1] * kv[0][0] * kv[0][1] * kv[0][2] ** 2 * kv[0][4] * kv[0][5] / kv[0][3] ** 2):
        if k == 6:      # stem: runs as 3x3 s1 on the space-to-depth image (12 -> 16 channels)
            ci2, k2, s2, h2, w2 = 16, 3, 1, h // 2, w // 2
        else:
            ci2, k2, s2, h2, w2 = ci, k, s, h, w
        co2 = (co + 15) // 16 * 16
        x = torch.randn(B, ci2, h2, w2, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wt = torch.randn(co2, ci2, k2, k2, device="cuda") / (ci2 * k2 * k2) ** 0.5
        y = T.conv_forward(x, wt, s2)
        dy = torch.randn_like(y)
        gf = 2.0 * B * y.shape[2] * y.shape[3] * co * ci * k * k / 1e9
        t_f = timeit(lambda: T.conv_forward(x, wt, s2))
        t_d = timeit(lambda: T.conv_dgrad(dy, wt, s2, h2, w2)) if k != 6 else 0.0
        t_w = timeit(lambda: T.conv_wgrad(x, dy, k2, s2))
        wb = wt.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        c_f = timeit(lambda: F.conv2d(x, wb, None, s2, k2 // 2))
        c_d = timeit(lambda: torch.ops.aten.convolution_backward(dy, x, wb, None, (s2, s2), (k2 // 2, k2 // 2), (1, 1), False, (0, 0), 1, (True, False, False))) if k != 6 else 0.0
        c_w = timeit(lambda: torch.ops.aten.convolution_backward(dy, x, wb, None, (s2, s2), (k2 // 2, k2 // 2), (1, 1), False, (0, 0), 1, (False, True, False)))
        tf = lambda t: gf / t * 1e3 if t else 0.0
        print(f"{ci:5d} {co:5d} {k} {s} {h:4d}x{w:<4d} {cnt:3d} | {gf:7.2f} | {t_f:7.1f} ({tf(t_f):5.0f}) | {t_d:7.1f} ({tf(t_d):5.0f}) | {t_w:7.1f} ({tf(t_w):5.0f}) | "
              f"{c_f:7.1f} / {c_d:7.1f} / {c_w:7.1f}", flush=True)
        for key, t in (("fwd", t_f), ("dgrad", t_d), ("wgrad", t_w), ("cudnn_fwd", c_f), ("cudnn_dgrad", c_d), ("cudnn_wgrad", c_w)):
            tot[key] += t * cnt
        tot_gf += gf * cnt
    print(f"per forward pass over {B} samples: {tot_gf:.1f} GFLOP")
    for key, t in tot.items():
        print(f"  {key:12s} {t / 1e3:8.2f} ms  {tot_gf / t * 1e3 if t else 0:7.1f} TF/s")


if __name__ == "__main__":
    main()
