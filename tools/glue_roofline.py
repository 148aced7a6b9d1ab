"""Training glue kernels (csrc/glue.cu) and the fused object loss (csrc/object_loss.cu) at YOLOPoint-L 640x640 batch-8 sizes:
CUDA-event time per call inside CUDA-graph replays (as the training step runs them; inputs rotated over > 2 x L2), algorithmic
bytes / time against the measured HBM copy bandwidth, and the same operation with the ATen kernels the reference's module tree runs.   python tools/glue_roofline.py > profiles/r02_glue_roofline.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from yolopoint_b200 import Model, losses as Lz, train as T  # noqa: E402

dev = torch.device("cuda")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(next((v for k, v in peaks.items() if "hbm" in k.lower() and isinstance(v, (int, float))), 6550.0))
CL = torch.channels_last


def timed(fn, n_sets, iters=30):
    """us per call, host launches (used for the eager object-loss comparison only)."""
    for i in range(3):
        fn(i % n_sets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_sets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def graph_timed(fn, n_sets, replays=10):
    """us per call with the calls of all input sets captured into ONE CUDA graph (no host launch cost between kernels)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(n_sets):
            fn(i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n_sets):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (replays * n_sets) * 1e3


def rnd(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16).contiguous(memory_format=CL)


A = torch.ops.aten
CODE = {"copy": 0, "up2": 1, "pool2": 2}
rows = []
B = 8
cases = [("cat(ups(xd), xc)  40x40", [(512, 20, "up2"), (512, 40, "copy")]), ("cat(ups(xe), xb)  80x80", [(256, 40, "up2"), (256, 80, "copy")]),
         ("C3 cat 160x160 (64+64)", [(64, 160, "copy"), (64, 160, "copy")]), ("C3 cat 80x80 (128+128)", [(128, 80, "copy"), (128, 80, "copy")]),
         ("v52 cat(pool(xa), ups(descB)) 80x80", [(128, 160, "pool2"), (128, 40, "up2")])]
with torch.no_grad():
    for name, parts in cases:
        modes = [m for _, _, m in parts]
        out_hw = parts[-1][1] * (2 if modes[-1] == "up2" else 1) // (2 if modes[-1] == "pool2" else 1)
        Ct = sum(c for c, _, _ in parts)
        nbytes_src = sum(B * c * hw * hw * 2 for c, hw, _ in parts)
        nbytes_out = B * Ct * out_hw * out_hw * 2
        n_sets = max(3, int(2 * 130e6 / (nbytes_src + nbytes_out)) + 1)
        sets = [[rnd(B, c, hw, hw) for c, hw, _ in parts] for _ in range(n_sets)]
        gouts = [rnd(B, Ct, out_hw, out_hw) for _ in range(n_sets)]
        mcodes = tuple(CODE[m] for m in modes)

        class Ctx:          # the backward launch of train._CatTC without the autograd engine around it
            needs_input_grad = (False,) + (True,) * len(parts)

            def save_for_backward(self, *t):
                self.saved_tensors = t
        ctxs = []
        for st in sets:
            c = Ctx()
            c.modes, c.geom, c.shapes = mcodes, (B, out_hw, out_hw), [tuple(t.shape) for t in st]
            c.saved_tensors = [t if m == "pool2" else None for t, m in zip(st, modes)]
            ctxs.append(c)

        def aten_fwd(i):
            outs = []
            for t, m in zip(sets[i], modes):
                outs.append(t if m == "copy" else F.interpolate(t, scale_factor=(2, 2), mode="nearest") if m == "up2" else F.max_pool2d(t, 2, 2))
            return torch.cat(outs, 1)
        idx = [[A.max_pool2d_with_indices(t, [2, 2], [2, 2])[1] if m == "pool2" else None for t, m in zip(st, modes)] for st in sets]

        def aten_bwd(i):
            res, c0 = [], 0
            for k, (t, m) in enumerate(zip(sets[i], modes)):
                gsl = gouts[i][:, c0:c0 + t.shape[1]]
                c0 += t.shape[1]
                if m == "copy":
                    res.append(gsl.contiguous(memory_format=CL))          # what the consuming kernel needs
                elif m == "up2":
                    res.append(A.upsample_nearest2d_backward(gsl, [out_hw, out_hw], list(t.shape), 2.0, 2.0))
                else:
                    res.append(A.max_pool2d_with_indices_backward(gsl, t, [2, 2], [2, 2], [0, 0], [1, 1], False, idx[i][k]))
            return res
        t_f = graph_timed(lambda i: T._CatTC.forward(Ctx(), mcodes, *sets[i]), n_sets)
        t_fa = graph_timed(aten_fwd, n_sets)
        t_b = graph_timed(lambda i: T._CatTC.backward(ctxs[i], gouts[i]), n_sets)
        t_ba = graph_timed(aten_bwd, n_sets)
        pool_extra = sum(B * c * hw * hw * 2 for c, hw, m in parts if m == "pool2")
        rows.append((name + " fwd", t_f, (nbytes_src + nbytes_out) / t_f / 1e3, t_fa))
        rows.append((name + " bwd", t_b, (nbytes_src + nbytes_out + pool_extra) / t_b / 1e3, t_ba))
        del sets, gouts, ctxs, idx

    for (C, hw) in ((512, 20), (256, 20)):
        n_sets = 24
        xs = [rnd(B, C, hw, hw) for _ in range(n_sets)]
        gs = [rnd(B, 4 * C, hw, hw) for _ in range(n_sets)]
        pool = lambda t: A.max_pool2d_with_indices(t, [5, 5], [1, 1], [2, 2])

        def aten_f(i):
            y1, i1 = pool(xs[i]); y2, i2 = pool(y1); y3, i3 = pool(y2)
            return torch.cat((xs[i], y1, y2, y3), 1), (y1, y2, i1, i2, i3)
        saved = [aten_f(i)[1] for i in range(n_sets)]

        def aten_b(i):
            y1, y2, i1, i2, i3 = saved[i]
            d0, d1, d2, d3 = gs[i].chunk(4, 1)
            pb = lambda g, x, ix: A.max_pool2d_with_indices_backward(g, x, [5, 5], [1, 1], [2, 2], [1, 1], False, ix)
            return d0 + pb(d1 + pb(d2 + pb(d3.contiguous(memory_format=CL), y2, i3), y1, i2), xs[i], i1)

        class SCtx:
            shape = (B, C, hw, hw)

            def save_for_backward(self, *t):
                self.saved_tensors = t
        sctx = [SCtx() for _ in range(n_sets)]
        for i in range(n_sets):
            T._SppfTC.forward(sctx[i], xs[i])
        t_f = graph_timed(lambda i: T._SppfTC.forward(SCtx(), xs[i]), n_sets)
        t_fa = graph_timed(lambda i: aten_f(i)[0], n_sets)
        t_b = graph_timed(lambda i: T._SppfTC.backward(sctx[i], gs[i]), n_sets)
        t_ba = graph_timed(aten_b, n_sets)
        nb = B * C * hw * hw * 2
        rows.append((f"SPPF pool cascade + cat {C}ch {hw}x{hw} fwd", t_f, (nb * 5 + nb * 3) / t_f / 1e3, t_fa))       # x, out4, arg maps
        rows.append((f"SPPF pool cascade + cat {C}ch {hw}x{hw} bwd", t_b, (nb * 4 + nb * 3 + nb) / t_b / 1e3, t_ba))

print("# Training glue kernels and fused object loss at YOLOPoint-L 640x640 batch-8 sizes (B200, CUDA events around CUDA-graph replays of the calls, inputs rotated over > 2 x L2)\n")
print(f"HBM peak used: {HBM:.0f} GB/s (MEASURED_PEAKS.json or the 6550 GB/s of round 1's records)\n")
print("| operation | csrc/glue.cu, us | algorithmic GB/s | % of HBM peak | ATen ops, us | speed-up |")
print("|---|---:|---:|---:|---:|---:|")
for name, t, gbs, ta in rows:
    print(f"| {name} | {t:.1f} | {gbs:.0f} | {100 * gbs / HBM:.0f} % | {ta:.1f} | {ta / t:.2f}x |")

# ---- object loss -------------------------------------------------------------------------------------------------------
torch.manual_seed(0)
mdl = Model(names=[str(i) for i in range(80)], version="l").to(dev)
cfg = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, iou_t=0.2, anchor_t=4.0, label_smoothing=0.0, fl_gamma=0.0)
crit = Lz.ComputeObjectLoss(mdl, cfg, dev)
gen = torch.Generator().manual_seed(1)
nt = 64
tg = torch.cat((torch.randint(0, B, (nt, 1), generator=gen).float(), torch.randint(0, 80, (nt, 1), generator=gen).float(),
                0.1 + 0.8 * torch.rand(nt, 2, generator=gen), 0.05 + 0.35 * torch.rand(nt, 2, generator=gen)), 1).to(dev)
p = [torch.randn(B, 3, s, s, 85, device=dev).requires_grad_(True) for s in (80, 40, 20)]
plan = crit.build_targets(p, tg)


def step(fused):
    crit.fused = fused
    loss, _ = crit(p, tg, plan)
    torch.autograd.grad(loss.sum(), p)


t_k = timed(lambda i: step(True), 1, iters=20)
t_a = timed(lambda i: step(False), 1, iters=20)
g1 = torch.cuda.CUDAGraph()
crit.fused = True
with torch.cuda.graph(g1):
    step(True)
t_g = timed(lambda i: g1.replay(), 1, iters=50)
cells = sum(B * 3 * s * s for s in (80, 40, 20))
print(f"\nObject loss forward + gradient, {nt} targets, {cells} cells x 85 logits (eager launches from Python, then the kernel path as one CUDA graph):\n")
print("| path | us per call |")
print("|---|---:|")
print(f"| csrc/object_loss.cu (3 memsets per level + claim / candidate / cells / finalize) | {t_k:.0f} |")
print(f"| the same replayed from a CUDA graph | {t_g:.0f} (d pred zero-fill + column write: {cells * 85 * 4 * 2 / t_g / 1e3:.0f} GB/s) |")
print(f"| PyTorch statement of the same loss (`ComputeObjectLoss._call_torch`, ~270 ATen launches forward + backward) | {t_a:.0f} |")
