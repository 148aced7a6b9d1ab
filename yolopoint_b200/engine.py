"""B200 inference engine for the YOLOPoint network (dataflow of src/models/YOLOPoint.py:198-246 of the reference).

The network is compiled into a flat launch list over preallocated NHWC activation buffers in HBM:

  * every ``Conv`` (conv + folded BN + SiLU) is one ``yp_conv2d_nhwc_fwd`` launch (tcgen05 implicit GEMM);
  * ``torch.cat`` never runs: producers store into channel slices of their consumer's buffer;
  * ``nn.Upsample(2,'nearest')`` never runs: the producer stores each pixel to the 2x2 block of the
    consumer's concat buffer (4 parity TMA stores) in its epilogue;
  * ``Bottleneck`` residual adds, the descriptor L2 normalisation and the C3 ``cv1 || cv2`` pair
    (one GEMM with N = 2c_) are fused into conv launches;
  * the three chained SPPF max-pools are one kernel writing the concat buffer in place.

Precision modes: ``fp32`` = activations/weights as (hi, lo) TF32 pairs, 3xTF32 tcgen05 MMAs, fp32-grade
results (the parity mode); ``bf16`` = bf16 operands, fp32 accumulation (the fast mode).

The launch list for a given (B, H, W) is captured into a CUDA graph after the first eager run.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import YP_ACT_NONE, YP_ACT_SILU, YP_ALGO_TCGEN05, YP_EPI_L2NORM, YP_FMT_BF16, YP_FMT_F32, YP_FMT_F32X2, YpChainOp, YpConvDesc, YpView

BN_EPS = 1e-3
MODEL_NAMES = ("YOLOPoint", "YOLOPointv52")
VERSIONS = {"n": (0.33, 0.25), "s": (0.33, 0.5), "m": (0.67, 0.75), "l": (1.0, 1.0), "x": (1.33, 1.25)}


def dims(version: str):
    import math
    dm, wm = VERSIONS[version]
    cs = tuple(math.ceil(2 ** k * wm / 8) * 8 for k in range(6, 11))
    ns = tuple(max(round(k * dm), 1) for k in (3, 6, 9))
    return cs, ns


def fold_conv_bn(w: torch.Tensor, bn) -> Tuple[torch.Tensor, torch.Tensor]:
    """W' = diag(g/sqrt(var+eps)) W, b' = beta - g*mu/sqrt(var+eps)   (src/utils/torch_utils_yolo.py:194-214)."""
    if isinstance(bn, dict):
        g, beta, mu, var, eps = bn["weight"], bn["bias"], bn["running_mean"], bn["running_var"], BN_EPS
    else:
        g, beta, mu, var, eps = bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, bn.eps
    scale = g.float().div(torch.sqrt(eps + var.float()))
    wf = w.float() * scale.view(-1, 1, 1, 1)
    bf = beta.float() - g.float().mul(mu.float()).div(torch.sqrt(var.float() + eps))
    return wf, bf


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """Round fp32 to TF32 (10-bit mantissa), ties away from zero == PTX cvt.rna.tf32.f32."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def split_tf32(x: torch.Tensor) -> torch.Tensor:
    """[...] fp32 -> [2, ...] (hi, lo) with hi = tf32(x), lo = tf32(x - hi)."""
    hi = tf32_round(x)
    return torch.stack((hi, tf32_round(x - hi)))


# --------------------------------------------------------------------------------------------------
# host-side description (CPU-testable: no CUDA needed to build it)
# --------------------------------------------------------------------------------------------------
@dataclass
class BufSpec:
    name: str
    H: int
    W: int
    C: int
    fmt: int


@dataclass
class SliceRef:
    buf: str
    c_off: int
    C: int
    upsample: int = 1


@dataclass
class ConvOp:
    names: Tuple[str, ...]      # state-dict prefixes whose outputs are concatenated along Cout
    src: SliceRef
    dst: Tuple[SliceRef, ...]
    k: int
    s: int
    cout: int                   # padded
    act: int
    bn: bool                    # Conv block (fold BN) vs bare nn.Conv2d
    residual: Optional[SliceRef] = None
    l2norm: bool = False
    stem: bool = False
    lane: int = 0               # 0 = trunk / detection branch, 1 = keypoint head, 2 = descriptor head, 3 / 4 = Detect levels 0 / 1


@dataclass
class PoolOp:
    buf: str


@dataclass
class Pool2Op:
    """MaxPool2d(2, 2) from one buffer slice into a channel slice of a half-resolution buffer (YOLOPointv52 descriptor head)."""
    src: SliceRef
    dst: SliceRef
    lane: int = 0


@dataclass
class L2NormOp:
    """desc / ||desc||_2 over the channels of every pixel of a fp32 buffer, in place (descriptor widths > 256: version "x")."""
    buf: str
    lane: int = 0


def _pad16(c):
    return (c + 15) // 16 * 16


class NetPlan:
    """Shape-independent part: op list with symbolic buffers; instantiate() gives the buffer extents."""

    def __init__(self, version: str, nc: int, precision: str = "fp32", model_name: str = "YOLOPoint"):
        assert precision in ("fp32", "bf16")
        assert model_name in MODEL_NAMES, model_name
        self.version, self.nc, self.no = version, nc, nc + 5
        self.model_name = model_name
        self.precision = precision
        self.act_fmt = YP_FMT_F32X2 if precision == "fp32" else YP_FMT_BF16
        (c1, c2, c3, c4, c5), (n1, n2, n3) = dims(version)
        self.c = (c1, c2, c3, c4, c5)
        self.D = c3
        self.bufs: Dict[str, Tuple[int, int, int]] = {}   # name -> (stride level, C, fmt)
        self.ops: List[object] = []
        self._lane = 0
        # the descriptor L2 normalisation runs in the epilogue of the last descriptor-head conv when all D channels of a pixel fit one
        # accumulator tile (<= 256 TMEM columns); version "x" (D = 320) runs that conv without it and yp_l2norm_nhwc behind it
        self.fused_l2norm = self.D <= 256
        self.det_pad = _pad16(3 * self.no)
        self.semi_pad = _pad16(65)
        (self._build if model_name == "YOLOPoint" else self._build_v52)(c1, c2, c3, c4, c5, n1, n2, n3)

    # level L means spatial (H / 2**L, W / 2**L)
    def _buf(self, name, level, C_, fmt=None):
        self.bufs[name] = (level, C_, self.act_fmt if fmt is None else fmt)
        return SliceRef(name, 0, C_)

    def _conv(self, names, src, dst, k, s, cout, act=True, bn=True, residual=None, l2norm=False, stem=False):
        if isinstance(names, str):
            names = (names,)
        if isinstance(dst, SliceRef):
            dst = (dst,)
        self.ops.append(ConvOp(tuple(names), src, tuple(dst), k, s, cout, YP_ACT_SILU if act else YP_ACT_NONE, bn, residual, l2norm, stem, self._lane))

    def _c3(self, name, src: SliceRef, level, cout, n, dst):
        c_ = cout // 2
        U = self._buf(name + ".U", level, 2 * c_)
        h = self._buf(name + ".h", level, c_)
        y = SliceRef(U.buf, 0, c_)
        self._conv((name + ".cv1", name + ".cv2"), src, U, 1, 1, 2 * c_)
        for i in range(n):
            self._conv(f"{name}.m.{i}.cv1", y, h, 1, 1, c_)
            self._conv(f"{name}.m.{i}.cv2", h, y, 3, 1, c_, residual=y)
        self._conv(name + ".cv3", U, dst, 1, 1, cout)

    def _build(self, c1, c2, c3, c4, c5, n1, n2, n3):
        S = SliceRef
        x0 = self._buf("in_s2d", 1, 16)
        t1 = self._buf("t1", 1, c1)
        t2 = self._buf("t2", 2, c2)
        xa = self._buf("xa", 2, c2)
        x3 = self._buf("x3", 3, c3)
        cat6 = self._buf("cat6", 3, 2 * c3)      # up(xe) | xb
        cat5 = self._buf("cat5", 4, 2 * c4)      # up(xd) | xc
        cat7 = self._buf("cat7", 4, 2 * c3)      # Conv8(xf) | xe
        cat8 = self._buf("cat8", 5, 2 * c4)      # Conv9(xg) | xd
        catd = self._buf("catd", 3, 2 * c2)      # ConvDescA(xa) | up(ConvDescB(xb))
        xb = S("cat6", c3, c3)
        xc = S("cat5", c4, c4)
        # shared encoder
        self._conv("Conv1", x0, t1, 3, 1, c1, stem=True)
        self._conv("Conv2", t1, t2, 3, 2, c2)
        self._c3("Bottleneck1", t2, 2, c2, n1, xa)
        self._conv("Conv3", xa, x3, 3, 2, c3)
        # keypoint head (lane 1: only needs x3)
        self._lane = 1
        sdet = self._buf("sdet", 3, c3)
        self._c3("BottleneckDet", x3, 3, c3, n1, sdet)
        semi = self._buf("semi", 3, self.semi_pad, YP_FMT_F32)
        self._conv("ConvDet", sdet, semi, 1, 1, self.semi_pad, act=False, bn=False)
        # desc + yolo encoder
        self._lane = 0
        self._c3("Bottleneck2", x3, 3, c3, n2, xb)
        # descriptor head (lane 2: needs xa and xb)
        self._lane = 2
        self._conv("ConvDescA", xa, S("catd", 0, c2), 3, 2, c2)
        self._conv("ConvDescB", xb, S("catd", c2, c2, upsample=2), 3, 2, c2)
        dd = self._buf("dd", 3, c3)
        self._c3("BottleneckDesc", catd, 3, c3, n1, dd)
        desc = self._buf("desc", 3, c3, YP_FMT_F32)
        self._conv("ConvDesc", dd, desc, 3, 1, c3, act=False, bn=False, l2norm=self.fused_l2norm)
        if not self.fused_l2norm:
            self.ops.append(L2NormOp("desc", self._lane))
        # yolo encoder
        self._lane = 0
        x4 = self._buf("x4", 4, c4)
        self._conv("Conv4", xb, x4, 3, 2, c4)
        self._c3("Bottleneck3", x4, 4, c4, n3, xc)
        x5 = self._buf("x5", 5, c5)
        self._conv("Conv5", xc, x5, 3, 2, c5)
        x5b = self._buf("x5b", 5, c5)
        self._c3("Bottleneck4", x5, 5, c5, n1, x5b)
        spp = self._buf("sppcat", 5, 2 * c5)     # x | y1 | y2 | y3, each c5/2
        self._conv("SPPooling.cv1", x5b, S("sppcat", 0, c5 // 2), 1, 1, c5 // 2)
        self.ops.append(PoolOp("sppcat"))
        x5c = self._buf("x5c", 5, c5)
        self._conv("SPPooling.cv2", spp, x5c, 1, 1, c5)
        # neck
        self._conv("Conv6", x5c, (S("cat8", c4, c4), S("cat5", 0, c4, upsample=2)), 1, 1, c4)     # xd
        x6 = self._buf("x6", 4, c4)
        self._c3("Bottleneck5", cat5, 4, c4, n1, x6)
        self._conv("Conv7", x6, (S("cat7", c3, c3), S("cat6", 0, c3, upsample=2)), 1, 1, c3)      # xe
        def detect_head(i, src, lvl, lane):
            # Detect.m[i] only needs its own pyramid level, so levels 0 and 1 are enqueued right after their input is produced.
            # YP_DETECT_LANES=1 moves them to side streams; measured on B200 it makes no difference (0.7697 vs 0.7685 ms per
            # network pass), so the default keeps them on the main lane
            det = self._buf(f"det{i}", lvl, self.det_pad, YP_FMT_F32)
            self._lane = lane if os.environ.get("YP_DETECT_LANES", "0") != "0" else 0
            self._conv(f"Detect.m.{i}", src, det, 1, 1, self.det_pad, act=False, bn=False)
            self._lane = 0

        xf = self._buf("xf", 3, c3)
        self._c3("Bottleneck6", cat6, 3, c3, n1, xf)
        detect_head(0, xf, 3, 3)
        self._conv("Conv8", xf, S("cat7", 0, c3), 3, 2, c3)
        xg = self._buf("xg", 4, c4)
        self._c3("Bottleneck7", cat7, 4, c4, n1, xg)
        detect_head(1, xg, 4, 4)
        self._conv("Conv9", xg, S("cat8", 0, c4), 3, 2, c4)
        xh = self._buf("xh", 5, c5)
        self._c3("Bottleneck8", cat8, 5, c5, n1, xh)
        detect_head(2, xh, 5, 0)

    def _c2f(self, name, src: SliceRef, level, cout, n, dst, cout_pad=None, l2norm=False):
        """C2f (src/models/common.py:151-165): cv1 1x1 -> 2c channels, chunk(2); every Bottleneckv8 (3x3 -> 3x3, no shortcut:
        C2f's default ``shortcut=False``) consumes the previous chunk and appends c channels; cv2 1x1 over all (2+n)c channels.
        ``chunk`` and ``cat`` never run: cv1 stores into channels [0, 2c) of one (2+n)c-channel buffer, bottleneck i reads
        slice [(1+i)c, (2+i)c) and stores into the next slice, cv2 reads the whole buffer."""
        c = int(cout * 0.5)
        assert c % 16 == 0, f"{name}: hidden width {c} is not a multiple of 16 channels"
        Y = self._buf(name + ".Y", level, (2 + n) * c)
        h = self._buf(name + ".h", level, c)
        self._conv(name + ".cv1", src, SliceRef(Y.buf, 0, 2 * c), 1, 1, 2 * c)
        for i in range(n):
            self._conv(f"{name}.m.{i}.cv1", SliceRef(Y.buf, (1 + i) * c, c), h, 3, 1, c)
            self._conv(f"{name}.m.{i}.cv2", h, SliceRef(Y.buf, (2 + i) * c, c), 3, 1, c)
        self._conv(name + ".cv2", Y, dst, 1, 1, cout if cout_pad is None else cout_pad, l2norm=l2norm)

    def _build_v52(self, c1, c2, c3, c4, c5, n1, n2, n3):
        """YOLOPointv52 (src/models/YOLOPoint.py:248-342): C2f blocks, no Conv6 / Conv7 / ConvDet / ConvDesc / ConvDescA; ``semi``
        and ``desc`` come straight out of a C2f (BN + SiLU, then the L2 normalisation for ``desc``); ``descA = MaxPool2d(2,2)(xa)``;
        SPPF on c4 channels; Detect on (c3, c4, c4)."""
        S = SliceRef
        x0 = self._buf("in_s2d", 1, 16)
        t1 = self._buf("t1", 1, c1)
        t2 = self._buf("t2", 2, c2)
        xa = self._buf("xa", 2, c2)
        x3 = self._buf("x3", 3, c3)
        cat6 = self._buf("cat6", 3, c4 + c3)     # up(xe) | xb
        cat5 = self._buf("cat5", 4, 2 * c4)      # up(xd) | xc
        cat7 = self._buf("cat7", 4, c3 + c4)     # Conv8(xf) | xe
        cat8 = self._buf("cat8", 5, 2 * c4)      # Conv9(xg) | xd
        catd = self._buf("catd", 3, 2 * c2)      # MaxPool(xa) | up(ConvDescB(xb))
        xb = S("cat6", c4, c3)
        xc = S("cat5", c4, c4)
        # shared encoder
        self._conv("Conv1", x0, t1, 3, 1, c1, stem=True)
        self._conv("Conv2", t1, t2, 3, 2, c2)
        self._c2f("Bottleneck1", t2, 2, c2, n1, xa)
        self._conv("Conv3", xa, x3, 3, 2, c3)
        # keypoint head (lane 1): semi = C2f(c3 -> 65) incl. BN + SiLU
        self._lane = 1
        semi = self._buf("semi", 3, self.semi_pad, YP_FMT_F32)
        self._c2f("BottleneckDet", x3, 3, 65, n1, semi, cout_pad=self.semi_pad)
        self._lane = 0
        self._c2f("Bottleneck2", x3, 3, c3, n2, xb)
        # descriptor head (lane 2)
        self._lane = 2
        self.ops.append(Pool2Op(xa, S("catd", 0, c2), self._lane))
        self._conv("ConvDescB", xb, S("catd", c2, c2, upsample=2), 3, 2, c2)
        desc = self._buf("desc", 3, c3, YP_FMT_F32)
        self._c2f("BottleneckDesc", catd, 3, c3, n1, desc, l2norm=self.fused_l2norm)
        if not self.fused_l2norm:
            self.ops.append(L2NormOp("desc", self._lane))
        # yolo encoder
        self._lane = 0
        x4 = self._buf("x4", 4, c4)
        self._conv("Conv4", xb, x4, 3, 2, c4)
        self._c2f("Bottleneck3", x4, 4, c4, n3, xc)
        x5 = self._buf("x5", 5, c4)
        self._conv("Conv5", xc, x5, 3, 2, c4)
        x5b = self._buf("x5b", 5, c4)
        self._c2f("Bottleneck4", x5, 5, c4, n1, x5b)
        spp = self._buf("sppcat", 5, 2 * c4)     # x | y1 | y2 | y3, each c4/2
        self._conv("SPPooling.cv1", x5b, S("sppcat", 0, c4 // 2), 1, 1, c4 // 2)
        self.ops.append(PoolOp("sppcat"))
        self._conv("SPPooling.cv2", spp, (S("cat8", c4, c4), S("cat5", 0, c4, upsample=2)), 1, 1, c4)          # xd
        # neck
        self._c2f("Bottleneck5", cat5, 4, c4, n1, (S("cat7", c3, c4), S("cat6", 0, c4, upsample=2)))           # xe

        def detect_head(i, src, lvl):
            det = self._buf(f"det{i}", lvl, self.det_pad, YP_FMT_F32)
            self._conv(f"Detect.m.{i}", src, det, 1, 1, self.det_pad, act=False, bn=False)

        xf = self._buf("xf", 3, c3)
        self._c2f("Bottleneck6", cat6, 3, c3, n1, xf)
        detect_head(0, xf, 3)
        self._conv("Conv8", xf, S("cat7", 0, c3), 3, 2, c3)
        xg = self._buf("xg", 4, c4)
        self._c2f("Bottleneck7", cat7, 4, c4, n1, xg)
        detect_head(1, xg, 4)
        self._conv("Conv9", xg, S("cat8", 0, c4), 3, 2, c4)
        xh = self._buf("xh", 5, c4)
        self._c2f("Bottleneck8", cat8, 5, c4, n1, xh)
        detect_head(2, xh, 5)

    def conv_ops(self):
        return [op for op in self.ops if isinstance(op, ConvOp)]

    def buffer_specs(self, H: int, W: int) -> List[BufSpec]:
        return [BufSpec(n, H >> l, W >> l, c, f) for n, (l, c, f) in self.bufs.items()]

    def flops_per_frame(self, H: int, W: int, weights_meta: Dict[str, Tuple[int, int, int]]) -> float:
        """Algorithmic conv FLOPs (unpadded channels, stem counted as the original 6x6)."""
        total = 0.0
        for op in self.conv_ops():
            lvl = self.bufs[op.dst[0].buf][0]
            ho, wo = H >> lvl, W >> lvl
            if op.dst[0].upsample == 2:
                ho, wo = ho // 2, wo // 2
            for n in op.names:
                co, ci, k = weights_meta[n]
                total += 2.0 * ho * wo * co * ci * k * k
        return total


# --------------------------------------------------------------------------------------------------
# layer chains: dependency analysis and segment order (host logic, CPU-testable)
# --------------------------------------------------------------------------------------------------
def _op_reads_writes(op) -> Tuple[List[Tuple[str, int, int]], List[Tuple[str, int, int]]]:
    """(reads, writes) of an op as (buffer, first channel, end channel) ranges."""
    if isinstance(op, (PoolOp, L2NormOp)):
        return [(op.buf, 0, 1 << 30)], [(op.buf, 0, 1 << 30)]
    if isinstance(op, Pool2Op):
        return [(op.src.buf, op.src.c_off, op.src.c_off + op.src.C)], [(op.dst.buf, op.dst.c_off, op.dst.c_off + op.dst.C)]
    reads = [(op.src.buf, op.src.c_off, op.src.c_off + op.src.C)]
    if op.residual is not None:
        reads.append((op.residual.buf, op.residual.c_off, op.residual.c_off + op.residual.C))
    return reads, [(d.buf, d.c_off, d.c_off + d.C) for d in op.dst]


def _overlap(a, b) -> bool:
    return any(x[0] == y[0] and x[1] < y[2] and y[1] < x[2] for x in a for y in b)


def chain_dependencies(ops: Sequence[object]) -> List[List[int]]:
    """For every op of a sequentially valid list: the earlier ops it must wait for -- read-after-write, write-after-read and
    write-after-write on channel ranges of the activation buffers -- transitively reduced (a dependency implied by another one is
    dropped).  This is what lets ``conv_chain_kernel`` run independent branches of the network (keypoint head, descriptor head,
    detection branch; src/models/YOLOPoint.py:198-246) concurrently and reorder them."""
    rw = [_op_reads_writes(op) for op in ops]
    anc: List[set] = []
    deps: List[List[int]] = []
    for j, (rj, wj) in enumerate(rw):
        direct = [i for i in range(j) if _overlap(rw[i][1], rj) or _overlap(rw[i][1], wj) or _overlap(rw[i][0], wj)]
        keep = [d for d in direct if not any(d in anc[e] for e in direct if e != d)]
        a = set(direct)
        for d in direct:
            a |= anc[d]
        anc.append(a)
        deps.append(keep)
    return deps


def chain_schedule(ops: Sequence[object], milestones: Sequence[int]) -> Tuple[List[int], List[int]]:
    """Order and segment boundaries for the chained execution of ``ops``.

    The order sorts the ops by dependency depth (ops of the same depth are independent: the branches of the network interleave, so
    consecutive items of the chain rarely wait for each other).  ``milestones`` are indices of ops after which work outside the chain
    may start (the last layer of a head, a Detect level): the sorted list is cut right behind each of them.
    Returns (order, cuts): ``order`` = op indices in execution order, ``cuts`` = segment end positions (exclusive) in ``order``."""
    deps = chain_dependencies(ops)
    depth = []
    for j, d in enumerate(deps):
        depth.append(1 + max((depth[i] for i in d), default=0))
    order = sorted(range(len(ops)), key=lambda j: (depth[j], j))
    pos = {j: k for k, j in enumerate(order)}
    cuts = sorted({pos[m] + 1 for m in milestones} | {len(order)})
    return order, cuts


def check_plan(net: "NetPlan", B: int, H: int, W: int) -> List[Tuple[str, str]]:
    """Dry run of every convolution launch of ``net`` at input shape [B,3,H,W] through the library's host-side planner
    (``yp_conv2d_plan_check``: formats, channel / tile geometry, shared memory and TMEM budgets).  Needs the shared library but no
    GPU and no weights: pointers are placeholders with the alignment real buffers have.  Returns ``[(layer, error message), ...]``
    (empty = every launch plans)."""
    L = _lib.lib()
    bad = []

    def view(ref: SliceRef) -> YpView:
        lvl, ctot, fmt = net.bufs[ref.buf]
        es = 2 if fmt == YP_FMT_BF16 else 4
        h, w = H >> lvl, W >> lvl
        v = YpView()
        v.base = 0x10000000 + ref.c_off * es
        v.B, v.C = B, ref.C
        v.H, v.W = (h // 2, w // 2) if ref.upsample == 2 else (h, w)
        v.pix_stride, v.plane_stride, v.format, v.upsample = ctot, B * h * w * ctot, fmt, ref.upsample
        return v

    for op in net.conv_ops():
        d = YpConvDesc()
        d.in_ = view(op.src)
        d.weight = 0x20000000
        d.ksize, d.stride, d.cout, d.act = op.k, op.s, op.cout, op.act
        d.epilogue = YP_EPI_L2NORM if op.l2norm else 0
        if op.residual is not None:
            d.residual = view(op.residual)
        d.n_out = len(op.dst)
        for i, ds in enumerate(op.dst):
            d.out[i] = view(ds)
        d.algo, d.split_k = YP_ALGO_TCGEN05, 0
        if L.yp_conv2d_plan_check(C.byref(d)) != 0:
            bad.append(("+".join(op.names), L.yp_last_error().decode("utf-8", "replace")))
    return bad


# --------------------------------------------------------------------------------------------------
# weight packing (pure torch, CPU or CUDA)
# --------------------------------------------------------------------------------------------------
def _get(sd, key):
    return sd[key] if key in sd else sd["model." + key]


def _has(sd, key):
    return key in sd or ("model." + key) in sd


def folded_weight(sd, name: str, bn: bool) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    if bn:
        w = _get(sd, name + ".conv.weight").float()
        if _has(sd, name + ".bn.weight"):
            return fold_conv_bn(w, {k: _get(sd, f"{name}.bn.{k}") for k in ("weight", "bias", "running_mean", "running_var")})
        return w, _get(sd, name + ".conv.bias").float()  # already fused state dict
    w = _get(sd, name + ".weight").float()
    b = _get(sd, name + ".bias").float() if _has(sd, name + ".bias") else None
    return w, b


def pack_conv(sd, op: ConvOp, cin_view: int, precision: str) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """-> (weight [planes, cout_pad, taps*cin_view], bias [cout_pad] or None) in the operand format."""
    ws, bs = [], []
    for n in op.names:
        w, b = folded_weight(sd, n, op.bn)
        ws.append(w)
        bs.append(b)
    w = torch.cat(ws, 0)
    bias = None if all(b is None for b in bs) else torch.cat([b if b is not None else torch.zeros(x.shape[0]) for b, x in zip(bs, ws)])
    co, ci, kh, kw = w.shape
    if op.stem:  # 6x6 s2 p2 on 3 channels == 3x3 s1 p1 on the 2x2 space-to-depth image (channel = (ph*2+pw)*3 + c)
        assert (kh, kw, ci) == (6, 6, 3)
        w = w.view(co, 3, 3, 2, 3, 2).permute(0, 2, 4, 3, 5, 1).reshape(co, 9, 12)  # [co, (dh,dw), (ph,pw,c)]
        wk = torch.zeros(co, 9, cin_view, dtype=torch.float32, device=w.device)
        wk[:, :, :12] = w
    else:
        assert ci <= cin_view and kh == op.k
        wk = torch.zeros(co, kh * kw, cin_view, dtype=torch.float32, device=w.device)
        wk[:, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
    full = torch.zeros(op.cout, wk.shape[1] * cin_view, dtype=torch.float32, device=w.device)
    full[:co] = wk.reshape(co, -1)
    if bias is not None:
        bfull = torch.zeros(op.cout, dtype=torch.float32, device=w.device)
        bfull[:co] = bias
        bias = bfull
    packed = split_tf32(full) if precision == "fp32" else full.to(torch.bfloat16).unsqueeze(0)
    return packed.contiguous(), bias


# --------------------------------------------------------------------------------------------------
# device side
# --------------------------------------------------------------------------------------------------
TUNING_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuning")


def tuning_path(version: str, B: int, H: int, W: int, precision: str, model_name: str = "YOLOPoint") -> str:
    tag = version if model_name == "YOLOPoint" else f"{model_name[len('YOLOPoint'):]}{version}"
    return os.path.join(TUNING_DIR, f"{tag}_{B}x{H}x{W}_{precision}.json")


def load_tuning(version: str, B: int, H: int, W: int, precision: str, model_name: str = "YOLOPoint") -> Dict[str, Tuple[int, int]]:
    """Per-layer (tile_n, split_k) measured on a B200 by tools/tune_conv.py; absent file -> library heuristics."""
    p = tuning_path(version, B, H, W, precision, model_name)
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        return {k: tuple(v) for k, v in json.load(f).get("layers", {}).items()}


_TORCH_DT = {YP_FMT_F32X2: torch.float32, YP_FMT_F32: torch.float32, YP_FMT_BF16: torch.bfloat16}
_PLANES = {YP_FMT_F32X2: 2, YP_FMT_F32: 1, YP_FMT_BF16: 1}


def make_view(t: torch.Tensor, fmt: int, c_off: int = 0, C_: Optional[int] = None, upsample: int = 1) -> YpView:
    """t: [planes, B, H, W, Ctot] contiguous device tensor."""
    P, B, H, W, Ct = t.shape
    C_ = Ct - c_off if C_ is None else C_
    v = YpView()
    v.base = t.data_ptr() + c_off * t.element_size()
    if upsample == 2:
        H, W = H // 2, W // 2
    v.B, v.H, v.W, v.C = B, H, W, C_
    v.pix_stride = Ct
    v.plane_stride = t.stride(0)
    v.format = fmt
    v.upsample = upsample
    return v


class ShapePlan:
    """Buffers + launch list for one (B, H, W)."""

    def __init__(self, eng: "Engine", B: int, H: int, W: int):
        assert H % 32 == 0 and W % 32 == 0, "H and W must be multiples of 32 (src/demo.py:112-121)"
        self.eng, self.B, self.H, self.W = eng, B, H, W
        net, dev = eng.net, eng.device
        self.bufs: Dict[str, torch.Tensor] = {}
        for s in net.buffer_specs(H, W):
            self.bufs[s.name] = torch.zeros((_PLANES[s.fmt], B, s.H, s.W, s.C), dtype=_TORCH_DT[s.fmt], device=dev)
        self.fmt = {n: f for n, (_, _, f) in net.bufs.items()}
        Hc, Wc = H // 8, W // 8
        self.x_in = torch.zeros((B, 3, H, W), dtype=torch.float32, device=dev)
        self.frame_in = torch.zeros((B, H, W, 3), dtype=torch.uint8, device=dev)
        self.semi = torch.empty((B, 65, Hc, Wc), dtype=torch.float32, device=dev)
        self.desc = torch.empty((B, net.D, Hc, Wc), dtype=torch.float32, device=dev)
        self.A = 3 * (Hc * Wc + (Hc // 2) * (Wc // 2) + (Hc // 4) * (Wc // 4))
        self.pred = torch.empty((B, self.A, net.no), dtype=torch.float32, device=dev)
        self.raw = [torch.empty((B, 3, Hc >> i, Wc >> i, net.no), dtype=torch.float32, device=dev) for i in range(3)]
        self._side = None
        self._keep = []      # ctypes objects referenced by the launch closures
        self.launches = []   # (lane, callable(stream_ptr) -> None, layer name)
        self._compile()
        self.graphs = {}

    def view(self, ref: SliceRef) -> YpView:
        return make_view(self.bufs[ref.buf], self.fmt[ref.buf], ref.c_off, ref.C, ref.upsample)

    def _compile(self):
        L = _lib.lib()
        eng = self.eng
        need, descs = {}, []
        self.conv_descs = []
        self.op_records = []   # per op of net.ops: ("conv", YpConvDesc) / ("pool", YpView) / ("pool2", None): input of the layer chains
        self.tuning = load_tuning(eng.net.version, self.B, self.H, self.W, eng.precision, eng.net.model_name) if eng.use_tuning else {}
        for op in eng.net.ops:
            if isinstance(op, PoolOp):
                v = self.view(SliceRef(op.buf, 0, self.bufs[op.buf].shape[-1]))
                self._keep.append(v)
                self.op_records.append(("pool", v))
                self.launches.append((0, lambda st, v=v: _lib.check(L.yp_sppf_pool(C.byref(v), st)), "sppf_pool"))
                continue
            if isinstance(op, L2NormOp):
                v = self.view(SliceRef(op.buf, 0, self.bufs[op.buf].shape[-1]))
                self._keep.append(v)
                self.op_records.append(("l2norm", v))
                self.launches.append((op.lane, lambda st, v=v: _lib.check(L.yp_l2norm_nhwc(C.byref(v), st)), "l2norm"))
                continue
            if isinstance(op, Pool2Op):
                vi, vo = self.view(op.src), self.view(op.dst)
                self._keep += [vi, vo]
                self.op_records.append(("pool2", None))
                self.launches.append((op.lane, lambda st, vi=vi, vo=vo: _lib.check(L.yp_maxpool2x2(C.byref(vi), C.byref(vo), st)), "maxpool2x2"))
                continue
            w, b = eng.weights[op.names]
            d = YpConvDesc()
            d.in_ = self.view(op.src)
            d.weight = w.data_ptr()
            d.bias = b.data_ptr() if b is not None else None
            d.ksize, d.stride, d.cout, d.act = op.k, op.s, op.cout, op.act
            d.epilogue = YP_EPI_L2NORM if op.l2norm else 0
            if op.residual is not None:
                d.residual = self.view(op.residual)
            d.n_out = len(op.dst)
            for i, ds in enumerate(op.dst):
                d.out[i] = self.view(ds)
            d.algo = eng.algo
            d.split_k = 0 if eng.split_k else 1
            tuned = self.tuning.get("+".join(op.names))
            if eng.tile_policy == "wide":
                # throughput plan (YP_TILE_WIDE): the widest N tile whose accumulator plan keeps the fp32-grade accuracy, K split only
                # where the accuracy bound asks for it -- a layer then occupies few SMs, which is what several frames in flight want
                if not op.l2norm:
                    d.tile_n = -max(1, int(eng.wide_grid_div))     # YP_TILE_WIDE, persistent grid capped at #SMs / wide_grid_div
                d.split_k = 0
            elif tuned and eng.split_k:
                d.tile_n, d.split_k = int(tuned[0]), int(tuned[1])
            self.conv_descs.append((op, d))
            self.op_records.append(("conv", d))
            need[op.lane] = max(need.get(op.lane, 0), int(L.yp_conv2d_workspace_bytes(C.byref(d)))) if eng.split_k else 0
            descs.append((op.lane, d))
            self._keep.append(d)
            self.launches.append((op.lane, lambda st, d=d: _lib.check(L.yp_conv2d_nhwc_fwd(C.byref(d), st)), "+".join(op.names)))
        # split-K scratch: one zero-initialised buffer per lane (lanes may run concurrently; launches of one lane are
        # stream-ordered and the arrival counters reset themselves)
        self.workspaces = {lane: torch.zeros(max(n, 16), dtype=torch.uint8, device=eng.device) for lane, n in need.items() if n > 0}
        for lane, d in descs:
            if lane in self.workspaces:
                d.workspace = self.workspaces[lane].data_ptr()
                d.workspace_bytes = self.workspaces[lane].numel()
        self.chain = None
        if eng.chain:
            self._build_chain()

    def _build_chain(self):
        """Layer chains (yp_conv_chain_*): the launch list reordered by dependency depth and cut behind the layers that work outside
        the network waits for (last layer of the keypoint / descriptor head, Detect levels 0 and 1); every segment is one persistent
        kernel.  Falls back to the per-layer launch list when an op cannot be chained (YOLOPointv52's 2x2 max pool, bf16 operands)."""
        L, ops = _lib.lib(), self.eng.net.ops
        if any(kind in ("pool2", "l2norm") for kind, _ in self.op_records) or self.eng.precision != "fp32" or self.eng.algo != YP_ALGO_TCGEN05:
            return
        names = ["+".join(op.names) if isinstance(op, ConvOp) else "sppf_pool" for op in ops]
        lanes = [getattr(op, "lane", 0) for op in ops]
        milestones = {}
        for lane in (1, 2):
            idx = [i for i, ln in enumerate(lanes) if ln == lane]
            if idx:
                milestones[max(idx)] = ("lane", lane)
        for nm in os.environ.get("YP_CHAIN_HOOK_CUTS", "Detect.m.0,Detect.m.1").split(","):
            if nm in names:
                milestones[names.index(nm)] = ("hook", nm)
        deps = chain_dependencies(ops)
        order, cuts = chain_schedule(ops, list(milestones))
        segs, k0 = [], 0
        for c in cuts:
            seg_ops = order[k0:c]
            local = {j: i for i, j in enumerate(seg_ops)}
            arr = (YpChainOp * len(seg_ops))()
            for i, j in enumerate(seg_ops):
                kind, rec = self.op_records[j]
                arr[i].type = 0 if kind == "conv" else 1
                if kind == "conv":
                    arr[i].conv = rec
                else:
                    arr[i].pool = rec
                dl = [local[d] for d in deps[j] if d in local]   # dependencies on earlier segments are ordered by the stream
                arr[i].n_deps = len(dl)
                for q, d in enumerate(dl):
                    arr[i].deps[q] = d
            handle = C.c_void_p()
            rc = L.yp_conv_chain_create(arr, len(seg_ops), C.byref(handle))
            if rc != 0:
                for sg in segs:
                    L.yp_conv_chain_destroy(sg["handle"])
                if os.environ.get("YP_CHAIN", "") == "require":
                    _lib.check(rc)
                return
            nk, ni, sm = C.c_int32(), C.c_int32(), C.c_int32()
            _lib.check(L.yp_conv_chain_info(handle, C.byref(nk), C.byref(ni), C.byref(sm)))
            segs.append({"handle": handle, "ops": seg_ops, "names": [names[j] for j in seg_ops], "kernels": nk.value, "items": ni.value, "smem": sm.value,
                         "lanes": [v for j in seg_ops if j in milestones for k_, v in [milestones[j]] if k_ == "lane"],
                         "hooks": [v for j in seg_ops if j in milestones for k_, v in [milestones[j]] if k_ == "hook"]})
            k0 = c
        self.chain = segs

    def __del__(self):
        try:
            if getattr(self, "chain", None):
                L = _lib.lib()
                for sg in self.chain:
                    L.yp_conv_chain_destroy(sg["handle"])
        except Exception:
            pass

    def n_net_launches(self) -> int:
        """Kernels one pass over the network launches (per-layer list, or the kernels of the layer chains)."""
        if self.chain:
            return sum(sg["kernels"] for sg in self.chain)
        return len(self.launches)

    # ---- pieces -------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.eng.device).cuda_stream)

    def run_input(self, frame: bool, src: Optional[torch.Tensor] = None):
        """Input conversion into the stem's operand buffer; ``src`` = another uint8 [B,H,W,3] device buffer than ``frame_in``."""
        L = _lib.lib()
        v = self.view(SliceRef("in_s2d", 0, 16))
        if frame:
            _lib.check(L.yp_frame_to_s2d((self.frame_in if src is None else src).data_ptr(), self.B, self.H, self.W, C.byref(v), self._stream()))
        else:
            _lib.check(L.yp_nchw_to_s2d(self.x_in.data_ptr(), self.B, self.H, self.W, C.byref(v), self._stream()))

    def side_stream(self, lane: int) -> "torch.cuda.Stream":
        if self._side is None:
            self._side = {ln: torch.cuda.Stream(self.eng.device) for ln in (1, 2, 3, 4)}
        return self._side[lane]

    def run_net(self, tails=None, after=None):
        """Launch list; the keypoint and descriptor heads run on side streams that fork from / join the current stream
        (also under graph capture), so their small grids overlap the detection branch.  ``tails`` maps a lane to a
        callable(stream_ptr) enqueued on that lane's stream after its last layer (e.g. heatmap + keypoint NMS).  ``after`` maps a
        layer name to a callable(stream_ptr) that only needs that layer's output: it is enqueued on its own side stream right
        behind the layer and joined at the end (e.g. the box-NMS candidate scan of a Detect level)."""
        tails = tails or {}
        after = after or {}
        dev = self.eng.device
        main = torch.cuda.current_stream(dev)
        if self.chain:
            return self._run_net_chain(tails, after, main)
        if not self.eng.multi_stream:
            st = C.c_void_p(main.cuda_stream)
            for _, f, name in self.launches:
                f(st)
                if name in after:
                    after[name](st)
            for lane in sorted(tails):
                tails[lane](st)
            return
        self.side_stream(1)
        started = {}
        hooks = []
        streams = {0: main}
        ptr = {0: C.c_void_p(main.cuda_stream)}
        free_hook_lanes = [ln for ln in (3, 4) if not any(l == ln for l, _, _ in self.launches)]
        for lane, f, name in self.launches:
            if lane and lane not in started:
                ev = torch.cuda.Event()
                ev.record(main)
                self._side[lane].wait_event(ev)
                started[lane] = True
                streams[lane] = self._side[lane]
                ptr[lane] = C.c_void_p(self._side[lane].cuda_stream)
            f(ptr[lane])
            if name in after:
                if free_hook_lanes:
                    hs = self._side[free_hook_lanes.pop(0)]
                    ev = torch.cuda.Event()
                    ev.record(streams[lane])
                    hs.wait_event(ev)
                    after[name](C.c_void_p(hs.cuda_stream))
                    hooks.append(hs)
                else:
                    after[name](ptr[lane])
        for lane in started:   # in the order the lanes started (a later lane's tail may wait for an earlier lane's event)
            if lane in tails:
                tails[lane](ptr[lane])
            ev = torch.cuda.Event()
            ev.record(self._side[lane])
            main.wait_event(ev)
        for hs in hooks:
            ev = torch.cuda.Event()
            ev.record(hs)
            main.wait_event(ev)

    def _run_net_chain(self, tails, after, main):
        """The network as layer chains: the segments run back to back on the current stream; a lane's tail / a layer's hook is
        enqueued on a side stream behind the segment that ends with that lane / layer, under the following segments."""
        L = _lib.lib()
        st = C.c_void_p(main.cuda_stream)
        multi = self.eng.multi_stream
        if multi:
            self.side_stream(1)
        joins = []
        free_hook_lanes = [3, 4]
        pending_hooks = dict(after)
        for sg in self.chain:
            _lib.check(L.yp_conv_chain_launch(sg["handle"], st))
            for lane in sorted(sg["lanes"]):
                if lane not in tails:
                    continue
                if not multi:
                    tails[lane](st)
                    continue
                side = self._side[lane]
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                tails[lane](C.c_void_p(side.cuda_stream))
                joins.append(side)
            for nm in sg["hooks"]:
                if nm not in pending_hooks:
                    continue
                hook = pending_hooks.pop(nm)
                if not multi or not free_hook_lanes:
                    hook(st)
                    continue
                hs = self._side[free_hook_lanes.pop(0)]
                ev = torch.cuda.Event()
                ev.record(main)
                hs.wait_event(ev)
                hook(C.c_void_p(hs.cuda_stream))
                joins.append(hs)
        for nm, hook in pending_hooks.items():   # hooks of layers that are not segment ends
            hook(st)
        for lane in sorted(tails):               # tails of lanes without a milestone (none in the shipped plans)
            if not any(lane in sg["lanes"] for sg in self.chain):
                tails[lane](st)
        for side in joins:
            ev = torch.cuda.Event()
            ev.record(side)
            main.wait_event(ev)

    def run_decode(self, want_raw: bool = True):
        L, net, st = _lib.lib(), self.eng.net, self._stream()
        row = 0
        for i in range(3):
            det = self.bufs[f"det{i}"]
            _, B, ny, nx, ldc = det.shape
            anc = (C.c_float * 6)(*[float(v) for v in self.eng.anchors_px[i]])
            _lib.check(L.yp_detect_decode(det.data_ptr(), B, ny, nx, ldc, 3, net.no, float(self.eng.stride[i]), anc,
                                          self.raw[i].data_ptr() if want_raw else None, self.pred.data_ptr(), self.A, row, st))
            row += 3 * ny * nx

    def run_export(self):
        """NHWC fp32 head outputs -> the NCHW tensors Model.forward returns."""
        L, st = _lib.lib(), self._stream()
        vs = self.view(SliceRef("semi", 0, self.bufs["semi"].shape[-1]))
        vd = self.view(SliceRef("desc", 0, self.bufs["desc"].shape[-1]))
        _lib.check(L.yp_nhwc_to_nchw(C.byref(vs), 65, self.semi.data_ptr(), st))
        _lib.check(L.yp_nhwc_to_nchw(C.byref(vd), self.eng.net.D, self.desc.data_ptr(), st))

    def run_forward(self, frame: bool = False):
        self.run_input(frame)
        self.run_net()
        self.run_decode(True)
        self.run_export()

    def graphed(self, key: str, fn):
        """Run ``fn`` through a CUDA graph captured on first use (after one eager warm-up run)."""
        g = self.graphs.get(key)
        if g is None:
            fn()  # eager: sets function attributes, fills caches
            torch.cuda.synchronize(self.eng.device)
            if not self.eng.use_graphs:
                return
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            self.graphs[key] = g
        g.replay()


class Engine:
    def __init__(self, sd, version: str, nc: int, device, precision: str = "fp32", algo: int = YP_ALGO_TCGEN05, use_graphs: bool = True,
                 multi_stream: bool = True, split_k: bool = True, use_tuning: bool = True, model_name: str = "YOLOPoint", chain: Optional[bool] = None,
                 tile_policy: Optional[str] = None, wide_grid_div: Optional[int] = None):
        _lib.lib(require_device=True)
        self.device = torch.device(device)
        self.net = NetPlan(version, nc, precision, model_name)
        self.precision, self.algo, self.use_graphs, self.multi_stream, self.split_k = precision, algo, use_graphs, multi_stream, split_k
        self.use_tuning = use_tuning
        # "latency" = per-layer (tile_n, split_k) from the measured tables (smallest time of ONE pass); "wide" = widest N tile, no split-K
        self.tile_policy = tile_policy or os.environ.get("YP_TILE_POLICY", "latency")
        assert self.tile_policy in ("latency", "wide"), self.tile_policy
        # wide plan: a layer's persistent kernel takes at most #SMs / wide_grid_div CTAs (1 = every SM; 3 measured best with 8 frames of
        # YOLOPoint-S in flight, where three layers of different frames then share the GPU)
        self.wide_grid_div = int(wide_grid_div or os.environ.get("YP_WIDE_GRID_DIV", "1"))
        # layer chains (one persistent kernel per network segment, yp_conv_chain_*): opt-in (chain=True or YP_CHAIN=1).  Measured on
        # B200 (YOLOPoint-S 640x640 batch 1, profiles/r02_chain.md): 1.005 ms per pass against 0.722 ms for the per-layer launch list
        # (CUDA graph + programmatic dependent launch) -- the per-layer cost is the CTA's own latency (first TMA, K loop, epilogue,
        # store drain), which a chain does not shorten, and multi-wave layers lose the two-CTAs-per-SM overlap.
        self.chain = (os.environ.get("YP_CHAIN", "0") not in ("0", "")) if chain is None else bool(chain)
        sd = {k: v.detach() for k, v in sd.items()}
        self.anchors = _get(sd, "Detect.anchors").float().cpu()
        self.stride = torch.tensor([8.0, 16.0, 32.0])
        self.anchors_px = (self.anchors * self.stride.view(-1, 1, 1)).reshape(3, 6).tolist()
        self.weights = {}
        for op in self.net.conv_ops():
            cin_view = op.src.C
            w, b = pack_conv({k: v.cpu() for k, v in sd.items() if any(k.startswith(n) or k.startswith("model." + n) for n in op.names)},
                             op, cin_view, precision)
            self.weights[op.names] = (w.to(self.device), None if b is None else b.to(self.device))
        self.plans: Dict[Tuple[int, int, int], ShapePlan] = {}

    def plan(self, B: int, H: int, W: int, slot: int = 0) -> ShapePlan:
        """Buffers + launch list for one input shape; `slot` selects an independent copy (own activation buffers, own
        graphs) so that several frames can be in flight on different streams."""
        key = (B, H, W, slot)
        if key not in self.plans:
            self.plans[key] = ShapePlan(self, B, H, W)
        return self.plans[key]

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """x: float32 NCHW [B,3,H,W] on this device -> the reference's eval-mode output dict (fresh tensors)."""
        if x.device != self.device:
            raise RuntimeError(f"input on {x.device}, engine on {self.device}; yolopoint_b200 has no CPU path")
        B, Cc, H, W = x.shape
        assert Cc == 3
        p = self.plan(B, H, W)
        p.x_in.copy_(x)
        p.graphed("forward", lambda: p.run_forward(False))
        return {"semi": p.semi.clone(), "desc": p.desc.clone(), "objects": (p.pred.clone(), [r.clone() for r in p.raw])}
