"""Deterministic synthetic weights / frames for parity tests and benchmarks.

There is no network in the build or benchmark environment, so no trained checkpoint exists
(reference weights are external downloads, README.md:33-37 of the reference).  Freshly initialised
weights give a flat keypoint heatmap and no boxes, which would leave the NMS / sampling / matching
stages idle.  ``perturb_state_dict`` rewrites a reference-format state dict in a seeded, documented
way so that every stage of the hot path does representative work:

  * BatchNorm affine + running statistics are randomised (exercises conv+BN folding,
    src/utils/torch_utils_yolo.py:194-214 of the reference),
  * every ``Conv`` weight gets a constant gain so activations stay O(1) with depth,
  * ``ConvDet.weight`` is scaled so the 65-way cell softmax is peaked (O(10^3) keypoints at 640x640),
  * ``Detect`` objectness / class biases are raised so O(10^2) boxes survive NMS.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch


# Per-version constants (conv_gain, head_gain, obj_bias, det_gain, cls_bias), tuned once on
# ``synthetic_frame(seed=0)`` so that the default thresholds (configs/kitti_inference.yaml:5-16 of the
# reference) select O(10^3) keypoints and O(10^2) boxes.  Purely elementwise on seeded values, hence
# bit-reproducible on any machine.
TUNING = {
    "n": (2.5, 32.0, -2.5, 12.0, -3.0),
    "s": (2.5, 12.0, -3.0, 12.0, -3.0),
    "m": (2.2, 12.0, -3.0, 12.0, -3.0),   # deeper nets need a smaller gain to stay O(1)
    "l": (2.0, 12.0, -3.0, 12.0, -3.0),
    "x": (2.0, 12.0, -3.0, 12.0, -3.0),
}


# YOLOPointv52 (no ConvDet: the keypoint logits are a BN + SiLU output, peaked by scaling BottleneckDet.cv2's BN weight): same
# columns, tuned the same way for the versions that have whole-frame fixtures / bench workloads (N, S); M / L keep the defaults.
V52_SEMI_GAIN = 8.0
TUNING_V52 = dict(TUNING, n=(2.5, 32.0, -3.3, 12.0, -3.0), s=(2.5, 40.0, -4.8, 12.0, -3.0))


def perturb_state_dict(sd: Dict[str, torch.Tensor], seed: int = 0, version: str = "s", conv_gain=None,
                       head_gain=None, obj_bias=None, det_gain=None, cls_bias=None, semi_gain=None) -> Dict[str, torch.Tensor]:
    """``semi_gain``: YOLOPointv52 has no ``ConvDet``; its keypoint logits are the BN + SiLU output of ``BottleneckDet.cv2``
    (src/models/YOLOPoint.py:283, 303), so the softmax is peaked by scaling that BN's affine weight instead (auto-detected
    from the keys when None; 1.0 for YOLOPoint)."""
    v52 = not any(k.endswith("ConvDet.weight") for k in sd)
    t = (TUNING_V52 if v52 else TUNING)[version]
    if semi_gain is None:
        semi_gain = V52_SEMI_GAIN if v52 else 1.0
    conv_gain = t[0] if conv_gain is None else conv_gain
    head_gain = t[1] if head_gain is None else head_gain
    obj_bias = t[2] if obj_bias is None else obj_bias
    det_gain = t[3] if det_gain is None else det_gain
    cls_bias = t[4] if cls_bias is None else cls_bias
    g = torch.Generator().manual_seed(1000 + seed)
    out = {}
    for k, v in sd.items():
        v = v.detach().clone()
        if k.endswith(".bn.weight"):
            v = torch.empty_like(v).uniform_(0.8, 1.2, generator=g)
            if semi_gain != 1.0 and k.endswith("BottleneckDet.cv2.bn.weight"):
                v = v * semi_gain
        elif k.endswith(".bn.bias"):
            v = torch.empty_like(v).normal_(0.0, 0.1, generator=g)
        elif k.endswith(".bn.running_mean"):
            v = torch.empty_like(v).normal_(0.0, 0.1, generator=g)
        elif k.endswith(".bn.running_var"):
            v = torch.empty_like(v).uniform_(0.5, 1.5, generator=g)
        elif k.endswith(".conv.weight"):
            v = v * conv_gain  # keeps activations O(1) through ~30 randomly initialised layers
        elif k.endswith("ConvDet.weight"):
            v = v * det_gain
        elif ".Detect.m." in k and k.endswith(".bias"):
            b = v.view(3, -1)
            b[:, 4] = obj_bias
            b[:, 5:] = cls_bias
            b[:, 5:] += torch.empty_like(b[:, 5:]).normal_(0.0, 1.0, generator=g)
            v = b.reshape(-1)
        elif ".Detect.m." in k and k.endswith(".weight"):
            v = v * head_gain
        out[k] = v
    return out


def synthetic_frame(H: int, W: int, seed: int = 0) -> np.ndarray:
    """uint8 [H,W,3] frame: smooth blobs + noise (more keypoint-like structure than white noise)."""
    rs = np.random.RandomState(seed)
    low = rs.randint(0, 256, (H // 16 + 1, W // 16 + 1, 3)).astype(np.float32)
    img = np.kron(low, np.ones((16, 16, 1), np.float32))[:H, :W]
    img = 0.6 * img + 0.4 * rs.randint(0, 256, (H, W, 3)).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)
