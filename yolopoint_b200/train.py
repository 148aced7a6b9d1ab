"""Training-step convolutions on the B200 kernels.

The reference trains by calling ``loss.backward()`` on the outputs of ``Model.forward`` in train mode
(src/train.py:208-220); autograd then needs, for every ``nn.Conv2d`` of src/models/common.py:22-34, the forward product,
the data gradient and the weight gradient.  Those three dense contractions are the training hot path (>97 % of the step's
FLOPs) and run here on the tcgen05 kernels of libyolopoint_b200.so:

  forward   y  = conv(x, W)              yp_conv2d_nhwc_fwd        (bf16 operands, fp32 accumulation, bf16 result)
  dgrad     dx = conv^T(dy, W)           yp_conv2d_nhwc_fwd on dy with transposed, tap-flipped weights; stride 2 = four
                                         launches with a custom tap list, one per output parity class (no zero insertion)
  wgrad     dW = x^T * dy                yp_conv2d_nhwc_wgrad      (MN-major tcgen05 GEMM over pixels, fp32 result)

Everything between the convolutions (BatchNorm statistics, SiLU, concat, upsample, max-pool, the losses, Adam) is PyTorch
glue on channels-last bf16 tensors, exactly the tensors the kernels read and write (no layout conversion in between).

``enable(model)`` switches a ``yolopoint_b200.Model`` in train mode to this path; there is no silent fallback: on a device
that is not sm_100 the call raises.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import YP_ACT_NONE, YP_ALGO_TCGEN05, YP_FMT_BF16, YP_UP_PARITY, YpConvDesc, YpView, YpWgradDesc

_CL = torch.channels_last


def _view(t: torch.Tensor, upsample: int = 1, hw=None) -> YpView:
    """t: [B,C,H,W] bf16 tensor that is contiguous in channels-last order -> NHWC view of all its channels."""
    B, Cc, H, W = t.shape
    assert t.dtype == torch.bfloat16 and t.is_contiguous(memory_format=_CL), (t.dtype, t.shape, t.stride())
    v = YpView()
    v.base = t.data_ptr()
    v.B, v.C = B, Cc
    v.H, v.W = hw if hw is not None else (H, W)
    v.pix_stride = Cc
    v.plane_stride = 0
    v.format = YP_FMT_BF16
    v.upsample = upsample
    return v


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cl(t: torch.Tensor) -> torch.Tensor:
    t = t if t.dtype == torch.bfloat16 else t.to(torch.bfloat16)
    return t.contiguous(memory_format=_CL)


def pack_weight(w: torch.Tensor) -> torch.Tensor:
    """[Co,Ci,kh,kw] -> bf16 [Co, taps*Ci] with K index = tap*Ci + ci (the layout of yp_conv2d_nhwc_fwd weights)."""
    co = w.shape[0]
    return w.permute(0, 2, 3, 1).reshape(co, -1).to(torch.bfloat16).contiguous()


# Packed operand copies of the master weights (nn.Parameter leaves only), reused until the optimizer changes the parameter (tensor
# version counter): a step uses every weight in two forward passes and two data-gradient passes.  An entry is only valid for the
# very parameter object it was made from (weak reference): a freed parameter's address may be reused by another one.
_PACK_CACHE: dict = {}


def _cached(kind: str, w: torch.Tensor, make):
    if not w.is_cuda or torch.cuda.is_current_stream_capturing() or not (w.is_leaf and isinstance(w, nn.Parameter)):
        # inside a CUDA graph the packing kernels must be part of the graph; temporaries (zero-padded / rearranged weights) get a
        # new allocation every call -- never cache those
        return make()
    key = (kind, id(w))
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0]() is w and hit[1] == w._version:
        return hit[2]
    out = make()
    if len(_PACK_CACHE) > 4096:
        _PACK_CACHE.clear()
    _PACK_CACHE[key] = (weakref.ref(w), w._version, out)
    return out


def _launch_conv(x, wp, y, ksize, stride, n_taps=0, dh=(), dw=(), up=1, out_hw=None):
    L = _lib.lib(require_device=True)
    d = YpConvDesc()
    d.in_ = _view(x)
    d.weight, d.bias = wp.data_ptr(), None
    d.ksize, d.stride, d.cout, d.act, d.epilogue, d.n_out = ksize, stride, wp.shape[0], YP_ACT_NONE, 0, 1
    d.out[0] = _view(y, upsample=up, hw=out_hw)
    d.algo, d.split_k = YP_ALGO_TCGEN05, 1
    if ksize == 0:
        d.n_taps = n_taps
        for i in range(n_taps):
            d.tap_dh[i], d.tap_dw[i] = dh[i], dw[i]
    _lib.check(L.yp_conv2d_nhwc_fwd(C.byref(d), _stream()))


def _dgrad_operands(w: torch.Tensor, stride: int):
    """Operand matrices of the data-gradient launches: stride 1 -> {"dgrad": [Ci, taps*Co]} (taps flipped, co/ci transposed);
    stride 2 -> one [Ci, n_taps*Co] matrix per output parity class."""
    Ci = w.shape[1]
    if stride == 1:
        return {"dgrad": w.flip(2, 3).permute(1, 2, 3, 0).reshape(Ci, -1).to(torch.bfloat16).contiguous()}
    out = {}
    for ph in range(2):
        for pw in range(2):
            taps = [(kh, kw) for kh in ([1] if ph == 0 else [0, 2]) for kw in ([1] if pw == 0 else [0, 2])]
            out[f"dgrad{ph}{pw}"] = torch.stack([w[:, :, kh, kw].t() for kh, kw in taps], 1).reshape(Ci, -1).to(torch.bfloat16).contiguous()
    return out


class WeightPack:
    """bf16 operand copies of one conv weight (forward + data-gradient layouts) in persistent buffers, refreshed once per optimizer
    step by ``refresh()`` -- outside the CUDA graphs of the forward / backward passes, which then contain no packing kernels (a
    step uses every weight in two forward and two data-gradient passes)."""

    def __init__(self, make_weight, stride: int, need_dgrad: bool, src: Optional[torch.Tensor] = None):
        self.make_weight, self.stride, self.need_dgrad, self.bufs = make_weight, stride, need_dgrad, None
        self.src, self.version = src, None      # the parameter the buffers were packed from and its version at that time

    def usable(self) -> bool:
        """True while the buffers hold the current weights.  Inside a captured graph the decision was taken at capture time (the
        replayed refresh keeps the buffers current); an eager call after an optimizer step sees a newer parameter version and packs
        on the fly instead of reading stale operands."""
        if self.bufs is None:
            return False
        return self.src is None or torch.cuda.is_current_stream_capturing() or self.src._version == self.version

    @torch.no_grad()
    def refresh(self):
        w = self.make_weight()
        if self.src is not None:
            self.version = self.src._version
        new = {"fwd": pack_weight(w)}
        if self.need_dgrad:
            new.update(_dgrad_operands(w, self.stride))
        if self.bufs is None:
            self.bufs = {k: v.clone() for k, v in new.items()}
        else:
            for k, v in new.items():
                self.bufs[k].copy_(v)


def conv_forward(x: torch.Tensor, w: torch.Tensor, stride: int, pack: Optional[WeightPack] = None) -> torch.Tensor:
    B, Ci, H, W = x.shape
    Co, _, k, _ = w.shape
    y = torch.empty((B, Co, H // stride, W // stride), dtype=torch.bfloat16, device=x.device, memory_format=_CL)
    wp = pack.bufs["fwd"] if pack is not None and pack.usable() else _cached("fwd", w, lambda: pack_weight(w))
    _launch_conv(x, wp, y, k, stride)
    return y


def conv_dgrad(dy: torch.Tensor, w: torch.Tensor, stride: int, H: int, W: int, pack: Optional[WeightPack] = None) -> torch.Tensor:
    """dx[b,ih,iw,ci] = sum_{kh,kw,co} dy[b,(ih+p-kh)/s,(iw+p-kw)/s,co] * w[co,ci,kh,kw]  (terms with integral indices)."""
    B, Co, Ho, Wo = dy.shape
    _, Ci, k, _ = w.shape
    dx = torch.empty((B, Ci, H, W), dtype=torch.bfloat16, device=dy.device, memory_format=_CL)
    if stride == 1:
        # a forward conv over dy with the taps flipped and (co, ci) transposed
        packed = pack.bufs if pack is not None and pack.usable() and "dgrad" in pack.bufs else None
        wt = packed["dgrad"] if packed else _cached("dgrad", w, lambda: _dgrad_operands(w, 1)["dgrad"])   # [Ci, taps*Co]
        _launch_conv(dy, wt, dx, k, 1)
        return dx
    assert k == 3 and stride == 2
    for ph in range(2):
        for pw in range(2):
            khs = [1] if ph == 0 else [0, 2]
            kws = [1] if pw == 0 else [0, 2]
            taps = [(kh, kw) for kh in khs for kw in kws]
            dh = [(ph + 1 - kh) // 2 for kh, _ in taps]
            dw = [(pw + 1 - kw) // 2 for _, kw in taps]
            key = f"dgrad{ph}{pw}"
            if pack is not None and pack.usable() and key in pack.bufs:
                wt = pack.bufs[key]
            else:
                wt = _cached(key, w, lambda: torch.stack([w[:, :, kh, kw].t() for kh, kw in taps], 1).reshape(Ci, -1)
                             .to(torch.bfloat16).contiguous())   # [Ci, n_taps*Co]
            _launch_conv(dy, wt, dx, 0, 1, len(taps), dh, dw, up=YP_UP_PARITY + 2 * ph + pw, out_hw=(Ho, Wo))
    return dx


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, ksize: int, stride: int) -> torch.Tensor:
    """-> fp32 [Co,Ci,k,k]"""
    L = _lib.lib(require_device=True)
    Ci, Co = x.shape[1], dy.shape[1]
    dwp = torch.zeros((Co, ksize * ksize * Ci), dtype=torch.float32, device=x.device)
    d = YpWgradDesc()
    d.x, d.dy = _view(x), _view(dy)
    d.ksize, d.stride, d.dw = ksize, stride, dwp.data_ptr()
    _lib.check(L.yp_conv2d_nhwc_wgrad(C.byref(d), _stream()))
    return dwp.view(Co, ksize, ksize, Ci).permute(0, 3, 1, 2)


class _Conv2dTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, pack=None):
        x = _cl(x)
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.pack = stride, pack
        return conv_forward(x, w, stride, pack)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _cl(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = conv_dgrad(dy, w, ctx.stride, x.shape[2], x.shape[3], ctx.pack)
        if ctx.needs_input_grad[1]:
            dw = conv_wgrad(x, dy, w.shape[2], ctx.stride).to(w.dtype)
        return dx, dw, None, None


def supported(w: torch.Tensor, stride: int, x: Optional[torch.Tensor] = None) -> bool:
    co, ci, k, k2 = w.shape
    ok = k == k2 and ((k == 1 and stride == 1) or (k == 3 and stride in (1, 2)))
    if x is not None:
        ok = ok and x.shape[2] % stride == 0 and x.shape[3] % stride == 0
    return ok


def _pad_weight(w: torch.Tensor) -> torch.Tensor:
    pad_o, pad_i = (-w.shape[0]) % 16, (-w.shape[1]) % 16
    return F.pad(w, (0, 0, 0, 0, 0, pad_i, 0, pad_o)) if (pad_o or pad_i) else w


def conv2d_tc(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, stride: int = 1, pack: Optional[WeightPack] = None) -> torch.Tensor:
    """Differentiable bias-free conv (pad = k // 2) on the tcgen05 kernels; x [B,Ci,H,W] (any float dtype, converted to
    channels-last bf16), w [Co,Ci,k,k] (fp32 master weights) -> bf16 channels-last [B,Co,H/s,W/s].  Channel counts that are
    not multiples of 16 (Detect: 255, ConvDet: 65) are zero-padded with differentiable torch ops around the kernel."""
    co, ci = w.shape[0], w.shape[1]
    pad_o, pad_i = (-co) % 16, (-ci) % 16
    if pad_i:
        x = F.pad(x, (0, 0, 0, 0, 0, pad_i))
    if pad_o or pad_i:
        w = F.pad(w, (0, 0, 0, 0, 0, pad_i, 0, pad_o))
    y = _Conv2dTC.apply(x, w, stride, pack)
    if pad_o:
        y = y[:, :co]
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1).to(y.dtype)
    return y


# ------------------------------------------------------------------------------------------------------------------
# BatchNorm (batch statistics) + SiLU of a Conv block
# ------------------------------------------------------------------------------------------------------------------
class _BnActTC(torch.autograd.Function):
    """act(bn(y)) of models/common.py:22-34 in train mode on yp_bn_act_fwd / yp_bn_act_bwd (bf16 NHWC, fp32 statistics)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, momentum, eps, act):
        L = _lib.lib(require_device=True)
        y = _cl(y)
        B, Cc, H, W = y.shape
        P = B * H * W
        out = torch.empty_like(y, memory_format=_CL)
        save = torch.empty(4 * Cc, dtype=torch.float32, device=y.device)
        acc = torch.empty(2 * Cc, dtype=torch.float32, device=y.device)
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        _lib.check(L.yp_bn_act_fwd(y.data_ptr(), P, Cc, g32.data_ptr(), b32.data_ptr(), running_mean.data_ptr() if running_mean is not None else None,
                                   running_var.data_ptr() if running_var is not None else None, float(momentum), float(eps), int(act), out.data_ptr(),
                                   save.data_ptr(), acc.data_ptr(), _stream()))
        ctx.save_for_backward(y, g32, save)
        ctx.act = int(act)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib(require_device=True)
        y, g32, save = ctx.saved_tensors
        dout = _cl(dout)
        B, Cc, H, W = y.shape
        dy = torch.empty_like(y, memory_format=_CL)
        gb = torch.empty(2 * Cc, dtype=torch.float32, device=y.device)
        _lib.check(L.yp_bn_act_bwd(dout.data_ptr(), y.data_ptr(), B * H * W, Cc, g32.data_ptr(), save.data_ptr(), ctx.act, dy.data_ptr(), gb.data_ptr(), _stream()))
        return dy, gb[Cc:], gb[:Cc], None, None, None, None, None


def bn_act_tc(y: torch.Tensor, bn: nn.BatchNorm2d, act: bool) -> torch.Tensor:
    out = _BnActTC.apply(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, act)
    if bn.num_batches_tracked is not None and not getattr(bn, "_yp_defer_count", False):
        bn.num_batches_tracked.add_(1)      # (trainer.TrainStep counts all BatchNorms of a step with one foreach add instead)
    return out


# ------------------------------------------------------------------------------------------------------------------
# glue between the convolutions: concat (+ nearest 2x upsampling / 2x2 max pooling of a part), SPPF pooling cascade
# ------------------------------------------------------------------------------------------------------------------
def _glue_ok(t: torch.Tensor) -> bool:
    return t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 4 and t.shape[1] % 8 == 0


class _CatTC(torch.autograd.Function):
    """torch.cat(parts, 1) of bf16 channels-last tensors where part i is first upsampled 2x (nearest) or 2x2 max-pooled according
    to ``modes[i]``: one launch forward (yp_cat_nhwc_fwd), one launch backward for all parts (yp_cat_nhwc_bwd)."""

    @staticmethod
    def forward(ctx, modes, *parts):
        L = _lib.lib(require_device=True)
        parts = [_cl(p) for p in parts]
        B = parts[0].shape[0]
        scale = {_lib.YP_CAT_COPY: (1, 1), _lib.YP_CAT_UP2: (2, 1), _lib.YP_CAT_POOL2: (1, 2)}
        hw = [(p.shape[2] * scale[m][0] // scale[m][1], p.shape[3] * scale[m][0] // scale[m][1]) for p, m in zip(parts, modes)]
        assert all(s == hw[0] for s in hw) and all(p.shape[0] == B for p in parts), [tuple(p.shape) for p in parts]
        H, W = hw[0]
        out = torch.empty((B, sum(p.shape[1] for p in parts), H, W), dtype=torch.bfloat16, device=parts[0].device, memory_format=_CL)
        arr = (_lib.YpCatPart * len(parts))()
        for i, (p, m) in enumerate(zip(parts, modes)):
            arr[i].src, arr[i].grad, arr[i].C, arr[i].mode = p.data_ptr(), None, p.shape[1], m
        _lib.check(L.yp_cat_nhwc_fwd(arr, len(parts), out.data_ptr(), B, H, W, _stream()))
        ctx.modes, ctx.geom = tuple(modes), (B, H, W)
        ctx.shapes = [tuple(p.shape) for p in parts]
        ctx.save_for_backward(*[p if m == _lib.YP_CAT_POOL2 else None for p, m in zip(parts, modes)])   # pooling routes by the source values
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib(require_device=True)
        dout = _cl(dout)
        B, H, W = ctx.geom
        n = len(ctx.modes)
        grads = [torch.empty(s, dtype=torch.bfloat16, device=dout.device, memory_format=_CL) if ctx.needs_input_grad[1 + i] else None
                 for i, s in enumerate(ctx.shapes)]
        arr = (_lib.YpCatPart * n)()
        for i in range(n):
            src = ctx.saved_tensors[i]
            arr[i].src = src.data_ptr() if src is not None else None
            arr[i].grad = grads[i].data_ptr() if grads[i] is not None else None
            arr[i].C, arr[i].mode = ctx.shapes[i][1], ctx.modes[i]
        _lib.check(L.yp_cat_nhwc_bwd(arr, n, dout.data_ptr(), B, H, W, _stream()))
        return (None, *grads)


def cat_tc(parts, modes=None) -> torch.Tensor:
    """cat over channels with per-part resampling (``modes``: "copy" | "up2" | "pool2"); the caller checked ``glue_ok``."""
    code = {"copy": _lib.YP_CAT_COPY, "up2": _lib.YP_CAT_UP2, "pool2": _lib.YP_CAT_POOL2}
    modes = [code[m] for m in (modes or ["copy"] * len(parts))]
    return _CatTC.apply(tuple(modes), *parts)


def glue_ok(parts) -> bool:
    """The concat / SPPF kernels take 1..4 bf16 CUDA tensors whose channel counts are multiples of 8."""
    return 1 <= len(parts) <= 4 and all(_glue_ok(p) for p in parts)


class _SppfTC(torch.autograd.Function):
    """cat(x, m(x), m(m(x)), m(m(m(x)))) with m = MaxPool2d(5, 1, 2): yp_sppf_train_fwd records the source pixel of every pooled
    value, yp_sppf_train_bwd scatters the three pooled gradients back to them."""

    @staticmethod
    def forward(ctx, x):
        L = _lib.lib(require_device=True)
        x = _cl(x)
        B, Cc, H, W = x.shape
        out = torch.empty((B, 4 * Cc, H, W), dtype=torch.bfloat16, device=x.device, memory_format=_CL)
        arg = torch.empty((3, B, H * W, Cc), dtype=torch.int16, device=x.device)
        _lib.check(L.yp_sppf_train_fwd(x.data_ptr(), out.data_ptr(), arg.data_ptr(), B, H, W, Cc, _stream()))
        ctx.save_for_backward(arg)
        ctx.shape = (B, Cc, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib(require_device=True)
        (arg,) = ctx.saved_tensors
        dout = _cl(dout)
        B, Cc, H, W = ctx.shape
        dx = torch.empty((B, Cc, H, W), dtype=torch.bfloat16, device=dout.device, memory_format=_CL)
        _lib.check(L.yp_sppf_train_bwd(dout.data_ptr(), arg.data_ptr(), dx.data_ptr(), B, H, W, Cc, _stream()))
        return dx


def sppf_cat_tc(x: torch.Tensor) -> torch.Tensor:
    return _SppfTC.apply(x)


# ------------------------------------------------------------------------------------------------------------------
# module-tree integration
# ------------------------------------------------------------------------------------------------------------------
def _stem_s2d(x: torch.Tensor, w: torch.Tensor):
    """6x6 s2 p2 conv on 3 channels == 3x3 s1 p1 conv on the 2x2 space-to-depth image (channel = (ph*2+pw)*3 + c); both
    rearrangements are differentiable torch ops, so autograd maps the kernel's [Co,12,3,3] gradient back to [Co,3,6,6]."""
    B, Cc, H, W = x.shape
    xs = x.view(B, Cc, H // 2, 2, W // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(B, 4 * Cc, H // 2, W // 2)
    co = w.shape[0]
    ws = w.view(co, Cc, 3, 2, 3, 2).permute(0, 3, 5, 1, 2, 4).reshape(co, 4 * Cc, 3, 3)
    return xs, ws


class TcConv2d(nn.Conv2d):
    """nn.Conv2d whose training-mode forward/backward run on the tcgen05 kernels (same parameters / state-dict keys)."""

    def forward(self, x):
        if not (self.training and x.is_cuda):
            return super().forward(x)
        k, s, p = self.kernel_size[0], self.stride[0], self.padding[0]
        w = self.weight
        if getattr(self, "_tc_cudnn", False):   # cross-check mode (tests): cuDNN on the same bf16 channels-last tensors
            return F.conv2d(_cl(x), w.to(torch.bfloat16), None if self.bias is None else self.bias.to(torch.bfloat16), self.stride, self.padding)
        pack = getattr(self, "_yp_pack", None)
        if (k, s, p) == (6, 2, 2) and w.shape[1] == 3 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
            xs, ws = _stem_s2d(x, w)
            return conv2d_tc(xs, ws, self.bias, 1, pack)
        if p == k // 2 and supported(w, s, x):
            return conv2d_tc(x, w, self.bias, s, pack)
        # geometry outside the kernels' range (not used by any YOLOPoint layer): cuDNN on the same bf16 channels-last tensors
        y = F.conv2d(_cl(x), w.to(torch.bfloat16), None if self.bias is None else self.bias.to(torch.bfloat16), self.stride, self.padding)
        return y


def _conv_block_forward(self, x):
    """Conv.forward (conv -> BN -> SiLU, models/common.py:30-31) with the BN + activation on the fused streaming kernels."""
    if self.training and x.is_cuda and fused_bn_block(self):
        return bn_act_tc(self.conv(x), self.bn, isinstance(self.act, nn.SiLU))
    return self._yp_plain_forward(x)


def fused_bn_block(block) -> bool:
    """True when a Conv block's BatchNorm + activation run on yp_bn_act_fwd / yp_bn_act_bwd in train mode."""
    bn = getattr(block, "bn", None)
    return (bn is not None and isinstance(block.conv, TcConv2d) and not block.conv._tc_cudnn and bn.momentum is not None
            and bn.track_running_stats and bn.weight is not None and bn.weight.shape[0] % 8 == 0 and isinstance(block.act, (nn.SiLU, nn.Identity)))


def _stem_weight(w: torch.Tensor) -> torch.Tensor:
    co, Cc = w.shape[0], w.shape[1]
    return w.view(co, Cc, 3, 2, 3, 2).permute(0, 3, 5, 1, 2, 4).reshape(co, 4 * Cc, 3, 3)


def attach_weight_packs(model: nn.Module):
    """Give every TcConv2d a WeightPack (persistent pre-packed operands) and return the list; call ``refresh()`` on each after
    every optimizer step (trainer.TrainStep does, from a CUDA graph).  The first convolution never needs a data gradient."""
    packs = []
    first = True
    for mod in model.modules():
        if isinstance(mod, TcConv2d):
            k, s, p = mod.kernel_size[0], mod.stride[0], mod.padding[0]
            stem = (k, s, p) == (6, 2, 2) and mod.weight.shape[1] == 3
            if not stem and not (p == k // 2 and supported(mod.weight, s)):
                continue
            make = (lambda m=mod: _pad_weight(_stem_weight(m.weight.detach()))) if stem else (lambda m=mod: _pad_weight(m.weight.detach()))
            mod._yp_pack = WeightPack(make, 1 if stem else s, need_dgrad=not (stem or first), src=mod.weight)
            packs.append(mod._yp_pack)
            first = False
    return packs


def detach_weight_packs(model: nn.Module):
    for mod in model.modules():
        if hasattr(mod, "_yp_pack"):
            del mod._yp_pack


def enable(model: nn.Module, cudnn_crosscheck: bool = False) -> nn.Module:
    """Route every nn.Conv2d of the module tree through the B200 kernels in train mode (in place; parameters are shared,
    state-dict keys unchanged).  BatchNorm keeps fp32 parameters / statistics and consumes the bf16 activations directly.
    ``cudnn_crosscheck`` keeps the same bf16 channels-last dataflow but calls cuDNN for the convolutions (test oracle)."""
    _lib.lib(require_device=True)
    from .model import Conv
    if not hasattr(Conv, "_yp_plain_forward"):
        Conv._yp_plain_forward = Conv.forward
        Conv.forward = _conv_block_forward
    for mod in model.modules():
        if type(mod) in (nn.Conv2d, TcConv2d):
            mod.__class__ = TcConv2d
            mod._tc_cudnn = cudnn_crosscheck
        mod._yp_glue = not cudnn_crosscheck         # concat / upsample / pooling on csrc/glue.cu (model._cat, SPPF.forward)
    model._tc_train = True
    return model


def disable(model: nn.Module) -> nn.Module:
    for mod in model.modules():
        if type(mod) is TcConv2d:
            mod.__class__ = nn.Conv2d
        mod._yp_glue = False
    model._tc_train = False
    return model
