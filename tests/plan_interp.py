"""CPU interpreter of the engine's launch plan (TEST INFRASTRUCTURE).

Executes yolopoint_b200.engine.NetPlan op by op with torch-CPU arithmetic, following the *contract* of the C ABI
(include/yolopoint_b200.h: channel-slice views, 2x-replicated stores, in-place residual, merged cv1||cv2, 2x2 max-pool into a concat slice, packed
weights incl. the space-to-depth stem, L2-norm epilogue, SPPF in the concat buffer).  It lets the CPU-only test
suite prove that the plan + weight packing reproduce the oracle network before any kernel runs on a GPU.
"""
import torch
import torch.nn.functional as F

from yolopoint_b200.engine import ConvOp, NetPlan, Pool2Op, PoolOp, pack_conv


def run_plan(net: NetPlan, sd, x: torch.Tensor, quantize=None):
    """x: [B,3,H,W] fp32 -> dict of NHWC buffers (value tensors [B,H,W,C], planes already summed)."""
    B, _, H, W = x.shape
    bufs = {s.name: torch.zeros(B, s.H, s.W, s.C) for s in net.buffer_specs(H, W)}
    # yp_nchw_to_s2d
    s2d = x.view(B, 3, H // 2, 2, W // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, H // 2, W // 2, 12)
    bufs["in_s2d"][..., :12] = s2d
    for op in net.ops:
        if isinstance(op, PoolOp):
            t = bufs[op.buf]
            c = t.shape[-1] // 4
            y = t[..., :c].permute(0, 3, 1, 2)
            for k in range(3):
                y = F.max_pool2d(y, 5, 1, 2)
                t[..., (k + 1) * c:(k + 2) * c] = y.permute(0, 2, 3, 1)
            continue
        if isinstance(op, Pool2Op):      # yp_maxpool2x2
            y = F.max_pool2d(bufs[op.src.buf][..., op.src.c_off:op.src.c_off + op.src.C].permute(0, 3, 1, 2), 2, 2)
            bufs[op.dst.buf][..., op.dst.c_off:op.dst.c_off + op.dst.C] = y.permute(0, 2, 3, 1)
            continue
        src = bufs[op.src.buf][..., op.src.c_off:op.src.c_off + op.src.C]
        w, b = pack_conv(sd, op, op.src.C, net.precision)
        wv = w.float().sum(0)                                          # [cout, taps*cin]
        k = 3 if op.stem else op.k
        wk = wv.view(op.cout, k, k, op.src.C).permute(0, 3, 1, 2)     # OIHW
        y = F.conv2d(src.permute(0, 3, 1, 2), wk, b, stride=op.s, padding=k // 2)
        if op.act:
            y = F.silu(y)
        y = y.permute(0, 2, 3, 1)
        if op.residual is not None:
            y = y + bufs[op.residual.buf][..., op.residual.c_off:op.residual.c_off + op.residual.C]
        if op.l2norm:
            y = y / torch.norm(y, p=2, dim=-1, keepdim=True)
        if quantize is not None:
            y = quantize(y)
        for d in op.dst:
            t = bufs[d.buf]
            if d.upsample == 2:
                t[..., d.c_off:d.c_off + d.C] = y.repeat_interleave(2, 1).repeat_interleave(2, 2)
            else:
                t[..., d.c_off:d.c_off + d.C] = y
    return bufs
