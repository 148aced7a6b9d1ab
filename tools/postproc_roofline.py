"""Achieved HBM bandwidth of the streaming (non-GEMM) kernels of the hot path at a size where bandwidth, not launch latency, is the
bound: BASELINE.json configs[2] geometry (1280x736 frames, batch 32 = the whole 8-GPU batch on one GPU unless --batch says otherwise).

  python tools/postproc_roofline.py [--batch 32] [--reps 20] [--out gpurun_out/postproc_roofline.json]

For every kernel: ALGORITHMIC bytes per launch (compulsory reads + writes) / the average launch time measured
with CUDA events on the launching stream, against the measured copy bandwidth of MEASURED_PEAKS.json.  Inputs rotate over enough
buffers (> 2 x L2 in total) that every launch reads from DRAM.  Prints one JSON line per kernel and a markdown table."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from yolopoint_b200 import _lib, ops  # noqa: E402
from yolopoint_b200._lib import YP_FMT_BF16, YP_FMT_F32, YP_FMT_F32X2  # noqa: E402
from yolopoint_b200.engine import make_view  # noqa: E402

L2_BYTES = 126e6


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return (json.load(open(p))["hbm_gbs"], "measured") if os.path.exists(p) else (6650.0, "fallback")


ONCE = False


def timed(fns, reps):
    """fns: list of callables (one per rotating buffer set); returns the mean time per launch in microseconds."""
    if ONCE:                     # one launch per kernel: the mode used under ncu (the time printed is not a measurement)
        fns[-1]()
        torch.cuda.synchronize()
        return float("nan")
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(reps, len(fns))
    e0.record()
    for i in range(n):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def n_sets(bytes_per_launch):
    return max(2, int(2.5 * L2_BYTES // max(1, bytes_per_launch)) + 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--H", type=int, default=736)
    ap.add_argument("--W", type=int, default=1280)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default="")
    ap.add_argument("--once", action="store_true", help="launch every kernel once, untimed (for ncu captures)")
    a = ap.parse_args()
    global ONCE
    ONCE = a.once
    L = _lib.lib(require_device=True)
    dev = torch.device("cuda")
    B, H, W = a.batch, a.H, a.W
    Hc, Wc = H // 8, W // 8
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    peak, src = peak_gbs()
    rows = []

    def report(name, what, nbytes, us):
        gbs = nbytes / (us * 1e-6) / 1e9
        r = {"kernel": name, "replaces": what, "batch": B, "algorithmic_MB": nbytes / 1e6, "us": us, "GBps": gbs, "frac_of_hbm_peak": gbs / peak,
             "peak_GBps": peak, "peak_source": src}
        rows.append(r)
        print(json.dumps(r), flush=True)

    # ---- input conversion: uint8 HWC frame -> space-to-depth NHWC operand (bf16 and (hi, lo) fp32)
    for fmt, tag, es, planes in ((YP_FMT_BF16, "bf16", 2, 1), (YP_FMT_F32X2, "f32x2", 4, 2)):
        nb = B * H * W * 3 + B * (H // 2) * (W // 2) * 16 * es * planes
        sets = []
        for _ in range(n_sets(nb)):
            fr = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
            out = torch.empty((planes, B, H // 2, W // 2, 16), dtype=torch.bfloat16 if es == 2 else torch.float32, device=dev)
            v = make_view(out, fmt)
            sets.append((fr, out, v))
        us = timed([lambda s=s: _lib.check(L.yp_frame_to_s2d(s[0].data_ptr(), B, H, W, C.byref(s[2]), st())) for s in sets], a.reps)
        report(f"to_s2d_kernel<frame> -> {tag}", "img/255 + HWC->CHW + stem re-layout", nb, us)
        del sets

    # ---- heatmap: 65-way cell softmax + depth-to-space, NHWC logits with the engine's 80-channel rows
    nb = B * Hc * Wc * 65 * 4 + B * H * W * 4
    sets = [(torch.randn(B, Hc, Wc, 80, device=dev) * 3, torch.empty(B, H, W, device=dev)) for _ in range(n_sets(nb))]
    us = timed([lambda s=s: ops.heatmap(s[0], layout="nhwc", out=s[1]) for s in sets], a.reps)
    report("heatmap_kernel (NHWC, 80-channel rows)", "softmax(65) + dustbin drop + PixelShuffle(8)", nb, us)
    del sets
    sets = [(torch.randn(B, 65, Hc, Wc, device=dev) * 3, torch.empty(B, H, W, device=dev)) for _ in range(n_sets(nb))]
    us = timed([lambda s=s: ops.heatmap(s[0], layout="nchw", out=s[1]) for s in sets], a.reps)
    report("heatmap_kernel (NCHW, flattenDetection API)", "softmax(65) + dustbin drop + PixelShuffle(8)", nb, us)
    del sets

    # ---- Detect decode (level 0 = 75 % of the rows): logits NHWC [B,ny,nx,256] -> raw + pred
    no, ny, nx = 85, Hc, Wc
    A0 = 3 * ny * nx
    anc = (C.c_float * 6)(10, 13, 16, 30, 33, 23)
    for want_raw in (True, False):
        nb = B * A0 * no * 4 * (3 if want_raw else 2)
        sets = []
        for _ in range(n_sets(nb)):
            det = torch.randn(B, ny, nx, 256, device=dev)
            raw = torch.empty(B, 3, ny, nx, no, device=dev) if want_raw else None
            pred = torch.empty(B, A0, no, device=dev)
            sets.append((det, raw, pred))
        us = timed([lambda s=s: _lib.check(L.yp_detect_decode(s[0].data_ptr(), B, ny, nx, 256, 3, no, 8.0, anc, s[1].data_ptr() if s[1] is not None else None,
                                                              s[2].data_ptr(), A0, 0, st())) for s in sets], a.reps)
        report("detect_decode_kernel" + (" (+raw copy)" if want_raw else ""), "sigmoid + grid/anchor decode + cat", nb, us)
        del sets

    # ---- box NMS front end on pred [B,A,85]: candidate scan reads the objectness column of every row (one 32-byte sector per
    # 340-byte row = the compulsory DRAM traffic), full rows only for the ~1 % that pass
    A = 3 * (Hc * Wc + (Hc // 2) * (Wc // 2) + (Hc // 4) * (Wc // 4))
    sets = []
    for _ in range(n_sets(B * A * no * 4)):
        pred = torch.rand(B, A, no, device=dev)
        pred[..., :2] *= 1000
        pred[..., 2:4] = pred[..., 2:4] * 60 + 10
        pred[..., 4] = (torch.rand(B, A, device=dev) < 0.01).float() * 0.9      # ~1 % of the rows are candidates
        pred[..., 5:] *= 0.3
        pred[..., 5] = 0.95                                                      # one label per candidate (conf 0.855)
        outb = (torch.zeros((B, 1000, 6), device=dev), torch.zeros((B,), dtype=torch.int32, device=dev))
        sets.append((pred, outb))
    us = timed([lambda s=s: ops.box_nms(s[0], 0.4, 0.45, True, True, 1000, cap=4096, out=s[1]) for s in sets], a.reps)
    cand = 0.01 * A
    nb = B * (A * 32 + cand * no * 4 + cand * cand / 8 + 1000 * 24)
    report("yp_box_nms (candidates + rank + mask + scan), ~1 % candidates", "non_max_suppression incl. torchvision.ops.nms", nb, us)
    assert int(sets[0][1][1].min()) >= 0, "candidate overflow"
    del sets

    # ---- descriptor sampling: N points per frame from the NHWC unit descriptors (D = 192: YOLOPoint-M)
    D, N = 192, 4096
    nb = B * N * D * 4 * 5 + B * N * 12
    sets = []
    for _ in range(n_sets(B * Hc * Wc * D * 4)):
        desc = torch.nn.functional.normalize(torch.randn(B, Hc, Wc, D, device=dev), dim=-1)
        pts = torch.stack((torch.randint(4, W - 4, (B, N), device=dev).float(), torch.randint(4, H - 4, (B, N), device=dev).float(),
                           torch.rand(B, N, device=dev)), -1).contiguous()
        cnt = torch.full((B,), N, dtype=torch.int32, device=dev)
        out = torch.empty(B, N, D, device=dev)
        sets.append((desc, pts, cnt, out))
    us = timed([lambda s=s: ops.sample_desc(s[0], s[1], s[2], (H, W), "nhwc", out=s[3]) for s in sets], a.reps)
    report(f"sample_desc_kernel (N={N}/frame, D={D})", "grid_sample + np.linalg.norm", nb, us)
    del sets

    # ---- 2x2 max pool (YOLOPointv52 descriptor head) on the stride-4 map, c2 = 96 channels (M)
    for fmt, tag, es, planes in ((YP_FMT_BF16, "bf16", 2, 1), (YP_FMT_F32X2, "f32x2", 4, 2)):
        Cc, H4, W4 = 96, H // 4, W // 4
        nb = B * H4 * W4 * Cc * es * planes * 1.25
        sets = []
        for _ in range(n_sets(nb)):
            dt = torch.bfloat16 if es == 2 else torch.float32
            src_t = torch.randn(planes, B, H4, W4, Cc, device=dev).to(dt)
            dst_t = torch.empty(planes, B, H4 // 2, W4 // 2, 2 * Cc, device=dev, dtype=dt)
            sets.append((src_t, dst_t, make_view(src_t, fmt), make_view(dst_t, fmt, 0, Cc)))
        us = timed([lambda s=s: _lib.check(L.yp_maxpool2x2(C.byref(s[2]), C.byref(s[3]), st())) for s in sets], a.reps)
        report(f"maxpool2x2_kernel ({tag})", "MaxPool2d(2,2) + cat", nb, us)
        del sets

    # ---- NHWC -> NCHW export of the descriptor map (Model.forward output)
    nb = B * Hc * Wc * D * 4 * 2
    sets = []
    for _ in range(n_sets(nb)):
        t = torch.randn(1, B, Hc, Wc, D, device=dev)
        sets.append((t, make_view(t, YP_FMT_F32), torch.empty(B, D, Hc, Wc, device=dev)))
    us = timed([lambda s=s: _lib.check(L.yp_nhwc_to_nchw(C.byref(s[1]), D, s[2].data_ptr(), st())) for s in sets], a.reps)
    report("nhwc_to_nchw_kernel (desc export)", "layout of Model.forward outputs", nb, us)
    del sets

    # ---- training: BatchNorm (batch statistics) + SiLU forward / backward on a YOLOPoint-L stride-2 activation (8 x 320 x 320 x 64)
    P, Cc = 8 * 320 * 320, 64
    sets = []
    for _ in range(n_sets(P * Cc * 2 * 2)):
        y = torch.randn(P, Cc, device=dev).to(torch.bfloat16)
        sets.append(dict(y=y, out=torch.empty_like(y), dout=torch.randn(P, Cc, device=dev).to(torch.bfloat16), dy=torch.empty_like(y),
                         g=torch.rand(Cc, device=dev) + 0.5, b=torch.randn(Cc, device=dev), rm=torch.zeros(Cc, device=dev), rv=torch.ones(Cc, device=dev),
                         save=torch.empty(4 * Cc, device=dev), acc=torch.empty(2 * Cc, device=dev), gb=torch.empty(2 * Cc, device=dev)))
    us = timed([lambda s=s: _lib.check(L.yp_bn_act_fwd(s["y"].data_ptr(), P, Cc, s["g"].data_ptr(), s["b"].data_ptr(), s["rm"].data_ptr(), s["rv"].data_ptr(),
                                                       0.03, 1e-3, 1, s["out"].data_ptr(), s["save"].data_ptr(), s["acc"].data_ptr(), st())) for s in sets], a.reps)
    report("yp_bn_act_fwd (statistics + normalise + SiLU)", "nn.BatchNorm2d (train) + nn.SiLU", P * Cc * 6, us)
    us = timed([lambda s=s: _lib.check(L.yp_bn_act_bwd(s["dout"].data_ptr(), s["y"].data_ptr(), P, Cc, s["g"].data_ptr(), s["save"].data_ptr(), 1, s["dy"].data_ptr(),
                                                       s["gb"].data_ptr(), st())) for s in sets], a.reps)
    report("yp_bn_act_bwd (reductions + input gradient)", "backward of BatchNorm2d + SiLU", P * Cc * 10, us)
    del sets

    print(f"\n| kernel | replaces | algorithmic MB / launch | us / launch | GB/s | % of HBM peak ({peak:.0f} GB/s, {src}) |\n|---|---|---:|---:|---:|---:|")
    for r in rows:
        print(f"| `{r['kernel']}` | {r['replaces']} | {r['algorithmic_MB']:.1f} | {r['us']:.1f} | {r['GBps']:.0f} | {100 * r['frac_of_hbm_peak']:.1f} |")
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
