"""Drop-in ``Model`` for the YOLOPoint hot path.

Mirrors the reference's model interface (src/models/YOLOPoint.py:17-145 ``Model``, :148-246 ``YOLOPoint``):
same constructor, same module tree -- hence identical ``state_dict()`` keys, ``parameters()`` order and
seeded initialisation -- and the same ``forward`` contract
``{'semi': [B,65,H/8,W/8], 'desc': [B,D,H/8,W/8] (unit norm), 'objects': (pred [B,A,nc+5], [raw_i])}``.

Execution:
  * eval mode  -> the B200 engine (``engine.Engine``): hand-written sm_100a kernels through the C ABI.  There is
    no fallback; a missing library or a non-sm_100 device raises.
  * train mode -> the same module tree under autograd; on a CUDA device every convolution (forward, data gradient, weight
    gradient: SURVEY.md section 8 row a11) runs on the tcgen05 kernels (``train.py``), bf16 operands / fp32 accumulation;
    BatchNorm, SiLU, concat, pooling and the losses stay PyTorch ops on the same channels-last tensors.
"""
from __future__ import annotations

import logging
import math
from copy import deepcopy
from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

LOGGER = logging.getLogger("yolopoint_b200")

ANCHORS_DEFAULT = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
VERSIONS = {"n": (0.33, 0.25), "s": (0.33, 0.5), "m": (0.67, 0.75), "l": (1.0, 1.0), "x": (1.33, 1.25)}
BN_EPS, BN_MOMENTUM = 1e-3, 0.03  # src/models/common.py:18-20


def make_divisible(x, divisor):  # src/utils/general_yolo.py:534
    return math.ceil(x / divisor) * divisor


def _cat(mod, parts, modes=None):
    """torch.cat(parts, 1) where part i is first passed through nn.Upsample(2, "nearest") (modes[i] == "up2") or nn.MaxPool2d(2, 2)
    ("pool2").  In train mode on the B200 backend (train.enable marks the modules) this is one kernel forward and one backward
    (csrc/glue.cu); otherwise the PyTorch ops of the reference."""
    modes = list(modes) if modes is not None else ["copy"] * len(parts)
    if getattr(mod, "_yp_glue", False) and mod.training:
        from . import train as _train
        if _train.glue_ok(parts):
            return _train.cat_tc(parts, modes)
    resample = {"copy": lambda t: t, "up2": lambda t: F.interpolate(t, scale_factor=(2, 2), mode="nearest"), "pool2": lambda t: F.max_pool2d(t, 2, 2)}
    return torch.cat([resample[m](t) for t, m in zip(parts, modes)], 1)


class Conv(nn.Module):
    """conv (no bias) -> BN -> SiLU (src/models/common.py:22-34)."""

    def __init__(self, c1, c2, k=1, s=1, p=None, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2 if p is None else p, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=BN_EPS, momentum=BN_MOMENTUM)
        self.act = nn.SiLU(inplace=True) if act else nn.Identity()

    def forward(self, x):
        return self.act(self.bn(self.conv(x))) if hasattr(self, "bn") else self.act(self.conv(x))


class Bottleneck(nn.Module):
    """x + cv2(cv1(x)): 1x1 then 3x3 (src/models/common.py:79-89)."""

    def __init__(self, c1, c2, shortcut=True, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_, c2, 3, 1)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C3(nn.Module):
    """cv3(cat(m(cv1(x)), cv2(x))) (src/models/common.py:123-135)."""

    def __init__(self, c1, c2, n=1, shortcut=True, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, e=1.0) for _ in range(n)))

    def forward(self, x):
        return self.cv3(_cat(self, (self.m(self.cv1(x)), self.cv2(x))))


class Bottleneckv8(nn.Module):
    """cv2(cv1(x)) with two 3x3 convs, optional shortcut (src/models/common.py:91-103)."""

    def __init__(self, c1, c2, shortcut=True, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    """cv2(cat(chunk(cv1(x), 2) + [m_i(previous)])) (src/models/common.py:151-165); ``shortcut`` defaults to False."""

    def __init__(self, c1, c2, n=1, shortcut=False, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneckv8(self.c, self.c, shortcut, k=(3, 3), e=1.0) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(_cat(self, y))


class SPPF(nn.Module):
    """src/models/common.py:213-229."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        x = self.cv1(x)
        if getattr(self, "_yp_glue", False) and self.training and self.m.kernel_size == 5:
            from . import train as _train
            if _train.glue_ok((x,)) and x.shape[2] * x.shape[3] <= 2048:
                return self.cv2(_train.sppf_cat_tc(x))      # pooling cascade + concat: one kernel forward, one backward (csrc/glue.cu)
        y1 = self.m(x)
        y2 = self.m(y1)
        return self.cv2(torch.cat((x, y1, y2, self.m(y2)), 1))


class Detect(nn.Module):
    """Detection head (src/models/yolo.py:34-91)."""
    stride = None

    def __init__(self, nc=80, anchors=(), ch=()):
        super().__init__()
        self.nc, self.no = nc, nc + 5
        self.nl, self.na = len(anchors), len(anchors[0]) // 2
        self.register_buffer("anchors", torch.tensor(anchors).float().view(self.nl, -1, 2))
        self.m = nn.ModuleList(nn.Conv2d(c, self.no * self.na, 1) for c in ch)

    def forward(self, xs):
        raw, z = [], []
        for i, x in enumerate(xs):
            x = self.m[i](x)
            bs, _, ny, nx = x.shape
            # reshape, not view: the B200 training path hands over channels-last tensors (bf16), returned to the losses as fp32
            x = x.float().reshape(bs, self.na, self.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
            raw.append(x)
            if not self.training:
                yv, xv = torch.meshgrid(torch.arange(ny, device=x.device), torch.arange(nx, device=x.device), indexing="ij")
                grid = torch.stack((xv, yv), 2).expand(1, self.na, ny, nx, 2).float()
                ag = (self.anchors[i] * self.stride[i]).view(1, self.na, 1, 1, 2)
                y = x.sigmoid()
                xy = (y[..., 0:2] * 2 - 0.5 + grid) * self.stride[i]
                wh = (y[..., 2:4] * 2) ** 2 * ag
                z.append(torch.cat((xy, wh, y[..., 4:]), -1).view(bs, -1, self.no))
        return raw if self.training else (torch.cat(z, 1), raw)


class YOLOPoint(nn.Module):
    """Module tree in the reference's construction order (src/models/YOLOPoint.py:148-196), so that seeded
    initialisation and ``state_dict`` keys coincide."""

    def __init__(self, width_multiple=1.0, depth_multiple=1.0, inp_ch=3, nc=80, anchors=None):
        super().__init__()
        c1, c2, c3, c4, c5 = [make_divisible(2 ** k * width_multiple, 8) for k in range(6, 11)]
        n1, n2, n3 = [max(round(k * depth_multiple), 1) for k in (3, 6, 9)]
        self.dims = (c1, c2, c3, c4, c5)
        self.depths = (n1, n2, n3)
        spec = [  # (name, factory) in reference order
            ("Conv1", lambda: Conv(inp_ch, c1, 6, 2, 2)), ("Conv2", lambda: Conv(c1, c2, 3, 2)),
            ("Bottleneck1", lambda: C3(c2, c2, n1)), ("Conv3", lambda: Conv(c2, c3, 3, 2)),
            ("Bottleneck2", lambda: C3(c3, c3, n2)), ("Conv4", lambda: Conv(c3, c4, 3, 2)),
            ("Bottleneck3", lambda: C3(c4, c4, n3)), ("Conv5", lambda: Conv(c4, c5, 3, 2)),
            ("Bottleneck4", lambda: C3(c5, c5, n1)), ("SPPooling", lambda: SPPF(c5, c5, 5)),
            ("Conv6", lambda: Conv(c5, c4, 1, 1, 0)), ("Bottleneck5", lambda: C3(c5, c4, n1)),
            ("Conv7", lambda: Conv(c4, c3, 1, 1, 0)), ("Bottleneck6", lambda: C3(c4, c3, n1)),
            ("Conv8", lambda: Conv(c3, c3, 3, 2, 1)), ("Bottleneck7", lambda: C3(c4, c4, n1)),
            ("Conv9", lambda: Conv(c4, c4, 3, 2, 1)), ("Bottleneck8", lambda: C3(c5, c5, n1)),
            ("Detect", lambda: Detect(nc, anchors=anchors, ch=(c3, c4, c5))),
            ("BottleneckDet", lambda: C3(c3, c3, n1)), ("ConvDet", lambda: nn.Conv2d(c3, 65, 1, 1, 0, bias=False)),
            ("ConvDescB", lambda: Conv(c3, c2, 3, 2, 1)), ("ConvDescA", lambda: Conv(c2, c2, 3, 2, 1)),
            ("ups", lambda: nn.Upsample(scale_factor=(2, 2), mode="nearest")),
            ("BottleneckDesc", lambda: C3(c3, c3, n1)), ("ConvDesc", lambda: nn.Conv2d(c3, c3, 3, 1, 1, bias=False)),
        ]
        for name, make in spec:
            setattr(self, name, make())

    def forward(self, x):  # PyTorch path (training); dataflow of src/models/YOLOPoint.py:198-246
        xa = self.Bottleneck1(self.Conv2(self.Conv1(x)))
        x = self.Conv3(xa)
        semi = self.ConvDet(self.BottleneckDet(x)).float()
        xb = self.Bottleneck2(x)
        desc = _cat(self, (self.ConvDescA(xa), self.ConvDescB(xb)), ("copy", "up2"))
        desc = self.ConvDesc(self.BottleneckDesc(desc)).float()
        desc = desc.div(torch.unsqueeze(torch.norm(desc, p=2, dim=1), 1))
        xc = self.Bottleneck3(self.Conv4(xb))
        xd = self.Conv6(self.SPPooling(self.Bottleneck4(self.Conv5(xc))))
        xe = self.Conv7(self.Bottleneck5(_cat(self, (xd, xc), ("up2", "copy"))))
        xf = self.Bottleneck6(_cat(self, (xe, xb), ("up2", "copy")))
        xg = self.Bottleneck7(_cat(self, (self.Conv8(xf), xe)))
        xh = self.Bottleneck8(_cat(self, (self.Conv9(xg), xd)))
        return {"semi": semi, "desc": desc, "objects": self.Detect([xf, xg, xh])}


class YOLOPointv52(nn.Module):
    """Module tree of the reference's YOLOPointv52 in its construction order (src/models/YOLOPoint.py:248-293): C2f blocks, no
    Conv6 / Conv7, ``semi`` and ``desc`` straight out of a C2f, ``descA = MaxPool2d(2, 2)(xa)``."""

    def __init__(self, width_multiple=1.0, depth_multiple=1.0, inp_ch=3, nc=80, anchors=None):
        super().__init__()
        c1, c2, c3, c4, c5 = [make_divisible(2 ** k * width_multiple, 8) for k in range(6, 11)]
        n1, n2, n3 = [max(round(k * depth_multiple), 1) for k in (3, 6, 9)]
        self.dims = (c1, c2, c3, c4, c5)
        self.depths = (n1, n2, n3)
        spec = [
            ("Conv1", lambda: Conv(inp_ch, c1, 6, 2, 2)), ("Conv2", lambda: Conv(c1, c2, 3, 2)),
            ("Bottleneck1", lambda: C2f(c2, c2, n1)), ("Conv3", lambda: Conv(c2, c3, 3, 2)),
            ("Bottleneck2", lambda: C2f(c3, c3, n2)), ("Conv4", lambda: Conv(c3, c4, 3, 2)),
            ("Bottleneck3", lambda: C2f(c4, c4, n3)), ("Conv5", lambda: Conv(c4, c4, 3, 2)),
            ("Bottleneck4", lambda: C2f(c4, c4, n1)), ("SPPooling", lambda: SPPF(c4, c4, 5)),
            ("Bottleneck5", lambda: C2f(c5, c4, n1)), ("Bottleneck6", lambda: C2f(c4 + c3, c3, n1)),
            ("Conv8", lambda: Conv(c3, c3, 3, 2, 1)), ("Bottleneck7", lambda: C2f(c4 + c3, c4, n1)),
            ("Conv9", lambda: Conv(c4, c4, 3, 2, 1)), ("Bottleneck8", lambda: C2f(c5, c4, n1)),
            ("Detect", lambda: Detect(nc, anchors=anchors, ch=(c3, c4, c4))),
            ("BottleneckDet", lambda: C2f(c3, 65, n1)), ("ConvDescB", lambda: Conv(c3, c2, 3, 2, 1)),
            ("MaxPool", lambda: nn.MaxPool2d(kernel_size=2, stride=2)),
            ("ups", lambda: nn.Upsample(scale_factor=(2, 2), mode="nearest")),
            ("BottleneckDesc", lambda: C2f(c3, c3, n1)),
        ]
        for name, make in spec:
            setattr(self, name, make())

    def forward(self, x):  # PyTorch path (training); dataflow of src/models/YOLOPoint.py:295-342
        xa = self.Bottleneck1(self.Conv2(self.Conv1(x)))
        x = self.Conv3(xa)
        semi = self.BottleneckDet(x).float()
        xb = self.Bottleneck2(x)
        desc = self.BottleneckDesc(_cat(self, (xa, self.ConvDescB(xb)), ("pool2", "up2"))).float()
        desc = desc.div(torch.unsqueeze(torch.norm(desc, p=2, dim=1), 1))
        xc = self.Bottleneck3(self.Conv4(xb))
        xd = self.SPPooling(self.Bottleneck4(self.Conv5(xc)))
        xe = self.Bottleneck5(_cat(self, (xd, xc), ("up2", "copy")))
        xf = self.Bottleneck6(_cat(self, (xe, xb), ("up2", "copy")))
        xg = self.Bottleneck7(_cat(self, (self.Conv8(xf), xe)))
        xh = self.Bottleneck8(_cat(self, (self.Conv9(xg), xd)))
        return {"semi": semi, "desc": desc, "objects": self.Detect([xf, xg, xh])}


_MODELS = {"YOLOPoint": YOLOPoint, "YOLOPointv52": YOLOPointv52}


class Model(nn.Module):
    """Same signature/behaviour as the reference wrapper (src/models/YOLOPoint.py:17-145)."""

    def __init__(self, names=(), model_name="YOLOPoint", version=None, inp_ch=3, anchors=None, precision="fp32"):
        super().__init__()
        anchors = anchors or ANCHORS_DEFAULT
        nc = len(names) if hasattr(names, "__len__") and len(names) > 0 else 1
        version = version.lower() if isinstance(version, str) else version
        if version not in VERSIONS:
            raise Exception(f"Version {version} is not a valid input. Choose one of n, s, m, l, x.")
        if model_name not in _MODELS:
            raise NotImplementedError(f"model_name={model_name!r}: only {sorted(_MODELS)} are on the accelerated hot path")
        if inp_ch != 3:
            raise NotImplementedError("the B200 stem kernel is specialised for 3 input channels")
        dm, wm = VERSIONS[version]
        self.version, self.nc, self.precision, self.model_name = version, nc, precision, model_name
        self.model = _MODELS[model_name](width_multiple=wm, depth_multiple=dm, inp_ch=inp_ch, nc=nc, anchors=anchors)
        m = self.model.Detect
        # The reference derives the strides with a 256x256 dummy forward in train mode (YOLOPoint.py:61-66).  The
        # result is always (8,16,32); its side effect -- one BN momentum update on an all-zero activation --
        # is reproduced exactly so that a seeded model has the same state dict.
        m.stride = torch.tensor([8.0, 16.0, 32.0])
        for mod in self.model.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.running_var.mul_(1 - BN_MOMENTUM)  # (1-m)*1 + m*0
                mod.num_batches_tracked.add_(1)
        m.anchors /= m.stride.view(-1, 1, 1)
        self._check_anchor_order(m)
        self._initialize_biases()
        self._engine = None
        self._fused = False
        # "b200": tcgen05 conv fwd/dgrad/wgrad on CUDA inputs; "torch": PyTorch/cuDNN fp32 autograd (the reference's own path);
        # "cudnn_bf16": the b200 dataflow (bf16 channels-last) with cuDNN convolutions -- cross-check for the tests
        self.train_backend = "b200"

    @staticmethod
    def _check_anchor_order(m):
        a = m.anchors.prod(-1).view(-1)
        if (a[-1] - a[0]).sign() != (m.stride[-1] - m.stride[0]).sign():
            LOGGER.info("Reversing anchor order")
            m.anchors[:] = m.anchors.flip(0)

    def _initialize_biases(self):  # YOLOPoint.py:92-100
        m = self.model.Detect
        for mi, s in zip(m.m, m.stride):
            b = mi.bias.view(m.na, -1)
            b.data[:, 4] += math.log(8 / (640 / s) ** 2)
            b.data[:, 5:] += math.log(0.6 / (m.nc - 0.999999))
            mi.bias = torch.nn.Parameter(b.view(-1), requires_grad=True)

    # -- execution ---------------------------------------------------------------------------
    def forward(self, x):
        if self.training:
            # CUDA tensors: the convolutions (forward, data and weight gradients) run on the tcgen05 kernels (train.py), bf16
            # operands with fp32 accumulation; CPU tensors: plain PyTorch autograd over the same parameters (host-side tests).
            if x.is_cuda and self.train_backend in ("b200", "cudnn_bf16"):
                if getattr(self, "_tc_train", None) != self.train_backend:
                    from . import train as _train
                    _train.enable(self, cudnn_crosscheck=self.train_backend == "cudnn_bf16")
                    self._tc_train = self.train_backend
            elif getattr(self, "_tc_train", None):
                from . import train as _train
                _train.disable(self)
                self._tc_train = None
            return self.model(x)
        return self.engine().forward(x)

    def engine(self):
        """The compiled B200 engine for the current weights (built lazily, dropped when weights change)."""
        if self._engine is None:
            from .engine import Engine
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("yolopoint_b200.Model runs inference on a CUDA (sm_100a) device only: call .cuda() first; "
                                   "there is no CPU fallback")
            self._engine = Engine(self.state_dict(), self.version, self.nc, dev, precision=self.precision, model_name=self.model_name,
                                  tile_policy=getattr(self, "tile_policy", None), wide_grid_div=getattr(self, "wide_grid_div", None))
        return self._engine

    def invalidate_engine(self):
        self._engine = None

    def train(self, mode=True):
        if mode:
            self._engine = None  # weights are about to change
        return super().train(mode)

    def _apply(self, fn):
        self = super()._apply(fn)
        m = self.model.Detect
        m.stride = fn(m.stride)
        self._engine = None
        return self

    def fuse(self):
        """Fold BN into the convs (YOLOPoint.py:84-90).  The engine always folds, so this only matters for the
        module tree / state dict, which are rewritten like the reference does."""
        from .engine import fold_conv_bn
        for mod in self.model.modules():
            if isinstance(mod, Conv) and hasattr(mod, "bn"):
                w, b = fold_conv_bn(mod.conv.weight.data, mod.bn)
                fused = nn.Conv2d(mod.conv.in_channels, mod.conv.out_channels, mod.conv.kernel_size, mod.conv.stride,
                                  mod.conv.padding, bias=True).requires_grad_(False).to(w.device)
                fused.weight.copy_(w)
                fused.bias.copy_(b)
                mod.conv = fused
                delattr(mod, "bn")
        self._fused = True
        self._engine = None
        return self

    def load_state_dict(self, target_state_dict, strict=True, verbose=False):  # YOLOPoint.py:102-119
        key = "model.Detect.m.0.bias" if "model.Detect.m.0.bias" in target_state_dict else "Detect.m.0.bias"
        self._engine = None
        if key in target_state_dict:
            if target_state_dict[key].shape == self.state_dict()[key].shape:
                return super().load_state_dict(target_state_dict, strict)
            if verbose:
                LOGGER.info("Number of classes have changed. Reinitializing Detect layer.\n")
            return self.load_partial_state_dict(target_state_dict, strict, verbose)
        try:
            return self.model.load_state_dict(target_state_dict, strict=strict)
        except RuntimeError:
            return super().load_state_dict(target_state_dict, strict=strict)

    def load_partial_state_dict(self, target_state_dict, strict=True, verbose=False):  # YOLOPoint.py:121-135
        current = self.state_dict()
        new = deepcopy(current)
        for k_cur, k_tgt in zip(current, target_state_dict):
            if ".".join(k_cur.split(".")[-2:]) == ".".join(k_tgt.split(".")[-2:]) and current[k_cur].shape == target_state_dict[k_tgt].shape:
                if verbose:
                    LOGGER.info(f"{k_cur} {' ' * (50 - len(k_cur))} {k_tgt}")
                new[k_tgt] = target_state_dict[k_cur]
        return super().load_state_dict(new, strict)

    def freeze_layers(self, to_freeze, verbose=True):  # YOLOPoint.py:137-145
        for i, (name, param) in enumerate(self.named_parameters()):
            if i in to_freeze:
                if verbose:
                    LOGGER.info(f"{i} {name} --> freeze")
                param.requires_grad = False


def load_model(meta_model=True, **kwargs):
    """src/utils/utils.py:55-57: ``meta_model=True`` (the way every reference script calls it) builds the ``Model`` wrapper from
    ``names / version / inp_ch / anchors / model_name``; ``meta_model=False`` builds the bare network class named by ``model_name``
    from ``width_multiple / depth_multiple / inp_ch / nc / anchors`` (what the wrapper itself does, src/models/YOLOPoint.py:47-53)."""
    if meta_model:
        return Model(**kwargs)
    name = kwargs.pop("model_name")
    if name not in _MODELS:
        raise NotImplementedError(f"model_name={name!r}: only {sorted(_MODELS)} are on the accelerated hot path")
    return _MODELS[name](**kwargs)
