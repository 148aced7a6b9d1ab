"""First-light diagnostics for the tcgen05 conv kernel (run on the GPU box).  Each case runs in its own process so that
a device trap in one does not poison the others.  Usage: python tools/gpu_diag.py            (driver, runs all cases)
                                                       python tools/gpu_diag.py CASE        (one case)"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = ["ident_1x1_f32", "ident_1x1_bf16", "bpattern_1x1_f32", "rand_1x1_f32", "rand_3x3_f32", "rand_3x3s2_f32", "rand_3x3_c16_f32",
         "rand_l2_f32", "rand_up_f32", "rand_3x3_bf16", "simt_3x3_f32"]


def one(case):
    import torch
    import torch.nn.functional as F
    from yolopoint_b200 import _lib
    from yolopoint_b200._lib import (YP_ACT_NONE, YP_ALGO_SIMT, YP_ALGO_TCGEN05, YP_FMT_BF16, YP_FMT_F32, YP_FMT_F32X2, YpConvDesc)
    from yolopoint_b200.engine import make_view, split_tf32
    import test_gpu_conv as T
    L = _lib.lib(require_device=True)
    dev = torch.device("cuda")
    print(f"== {case}: device {torch.cuda.get_device_name(0)}", flush=True)
    if case.startswith("ident") or case.startswith("bpattern"):
        fmt = YP_FMT_BF16 if case.endswith("bf16") else YP_FMT_F32X2
        B, H, W, Cc = 1, 16, 16, 32
        if case.startswith("ident"):
            pix = torch.arange(H * W, dtype=torch.float32).view(1, H, W, 1)
            ch = torch.arange(Cc, dtype=torch.float32).view(1, 1, 1, Cc)
            x = (pix + ch / 64.0).to(dev)         # exactly representable in bf16? pix<256 (8 bits) + ch/64 -> needs 14 bits: use smaller
            if fmt == YP_FMT_BF16:
                x = ((pix % 16) * 4 + ch / 32.0 * 0 + (ch % 4)).to(dev)
            w = torch.eye(Cc, device=dev)
        else:
            x = torch.ones(B, H, W, Cc, device=dev)
            w = (torch.arange(Cc, dtype=torch.float32).view(Cc, 1) * 64 + torch.arange(Cc, dtype=torch.float32).view(1, Cc)).to(dev) / 4096.0
        in_buf, x_eff = T._mk_act(x, fmt, Cc, 0)
        wp = w.to(torch.bfloat16).unsqueeze(0).contiguous() if fmt == YP_FMT_BF16 else split_tf32(w).contiguous()
        out_fmt = YP_FMT_F32
        out_buf = torch.full((1, B, H, W, Cc), -1.0, device=dev)
        d = YpConvDesc()
        d.in_ = make_view(in_buf, fmt, 0, Cc)
        d.weight, d.bias = wp.data_ptr(), None
        d.ksize, d.stride, d.cout, d.act, d.epilogue, d.n_out = 1, 1, Cc, YP_ACT_NONE, 0, 1
        d.out[0] = make_view(out_buf, out_fmt, 0, Cc)
        d.algo = YP_ALGO_TCGEN05
        rc = L.yp_conv2d_nhwc_fwd(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        print("rc", rc, L.yp_last_error())
        torch.cuda.synchronize()
        w_eff = wp.float().sum(0)
        exp = torch.einsum("bhwc,oc->bhwo", x_eff, w_eff)
        got = out_buf[0]
        diff = (got - exp).abs()
        print("max abs diff", float(diff.max()), "mismatching elems", int((diff > 1e-3).sum()), "of", diff.numel())
        if float(diff.max()) > 1e-3:
            g2, e2 = got.view(H * W, Cc).cpu(), exp.view(H * W, Cc).cpu()
            torch.set_printoptions(linewidth=250, precision=3, sci_mode=False)
            for r in (0, 1, 2, 7, 8, 9, 17, 255):
                print(f"row {r} got", g2[r, :16].tolist())
                print(f"row {r} exp", e2[r, :16].tolist())
            bad_rows = (diff.view(H * W, Cc) > 1e-3).any(1).nonzero().flatten().tolist()
            bad_cols = (diff.view(H * W, Cc) > 1e-3).any(0).nonzero().flatten().tolist()
            print("bad rows", bad_rows[:64], "... total", len(bad_rows))
            print("bad cols", bad_cols)
        return
    table = {
        "rand_1x1_f32": (T.CASES[0], YP_FMT_F32X2, YP_ALGO_TCGEN05), "rand_3x3_f32": (T.CASES[1], YP_FMT_F32X2, YP_ALGO_TCGEN05),
        "rand_3x3s2_f32": (T.CASES[2], YP_FMT_F32X2, YP_ALGO_TCGEN05), "rand_3x3_c16_f32": (T.CASES[5], YP_FMT_F32X2, YP_ALGO_TCGEN05),
        "rand_l2_f32": (T.CASES[3], YP_FMT_F32X2, YP_ALGO_TCGEN05), "rand_up_f32": (T.CASES[4], YP_FMT_F32X2, YP_ALGO_TCGEN05),
        "rand_3x3_bf16": (T.CASES[1], YP_FMT_BF16, YP_ALGO_TCGEN05), "simt_3x3_f32": (T.CASES[1], YP_FMT_F32X2, YP_ALGO_SIMT),
    }
    c, fmt, algo = table[case]
    try:
        err = T.run_case(c, fmt, algo)
        print("PASS rel err", err)
    except AssertionError as e:
        print("FAIL", e)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1])
    else:
        for c in CASES:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), c], capture_output=True, text=True, timeout=300)
            print(r.stdout[-4000:])
            if r.returncode != 0:
                print(f"-- {c} exited with {r.returncode}\n{r.stderr[-3000:]}")
