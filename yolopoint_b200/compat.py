"""Executable form of the import switch of INTEGRATION.md: patch the reference's own modules so that its scripts run unchanged.

    import sys; sys.path.insert(0, "<reference>/src")
    import yolopoint_b200.compat as compat
    compat.install()              # before the reference scripts build their model / frontend
    import demo                   # src/demo.py: YoloPointFrontend now builds yolopoint_b200.Model, post-processing runs on the kernels

The reference has no plugin registry; its seam is the set of names its scripts resolve at run time:
``load_model`` resolves ``models.Model`` through ``getattr(import_module('models'), ...)`` (src/utils/utils.py:55-57), and the
post-processing functions are module attributes of ``utils.utils`` / ``utils.general_yolo`` / ``evaluations.descriptor_evaluation`` that
``demo.py`` / ``train.py`` / the export scripts import by name.  ``install()`` rebinds exactly those attributes; ``uninstall()``
restores them.  Nothing here touches the reference's files.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, List, Tuple

_SAVED: List[Tuple[object, str, object]] = []

# reference module -> {attribute: name in yolopoint_b200}
BINDINGS: Dict[str, Dict[str, str]] = {
    "models": {"Model": "Model"},
    "models.YOLOPoint": {"Model": "Model"},
    "utils.utils": {"flattenDetection": "flattenDetection", "getPtsFromHeatmap": "getPtsFromHeatmap", "getPtsFromSemi": "getPtsFromSemi",
                    "nms_fast": "nms_fast"},
    "utils.general_yolo": {"non_max_suppression": "non_max_suppression"},
    "evaluations.descriptor_evaluation": {"sample_desc_from_points": "sample_desc_from_points"},
    "models.model_wrap": {"PointTracker": "PointTracker"},
    "demo": {"PointTracker": "PointTracker", "non_max_suppression": "non_max_suppression", "nms_fast": "nms_fast"},
}


# training-side names (src/train.py:7-8, 12-13): opt-in, the inference switch does not need them
LOSS_BINDINGS: Dict[str, Dict[str, str]] = {
    "utils.loss_functions": {"ComputeObjectLoss": "losses.ComputeObjectLoss", "ComputeDetectorLoss": "losses.ComputeDetectorLoss",
                             "infonce": "losses.infonce", "descriptor_loss_sparse": "losses.descriptor_loss_sparse"},
    "utils.utils": {"labels2Dto3D": "losses.labels2Dto3D", "getMasks": "losses.getMasks"},
}


def _resolve(root, dotted: str):
    obj = root
    for part in dotted.split("."):
        obj = getattr(obj, part)
    return obj


def install(modules=None, losses: bool = False) -> List[str]:
    """Rebind the hot-path names of the (importable) reference modules to their yolopoint_b200 counterparts.  Modules that cannot be
    imported in this environment (missing optional dependencies of the reference) are skipped.  ``losses=True`` also rebinds the
    training losses / target helpers.  Returns the rebound names."""
    import yolopoint_b200 as yp
    from . import losses as _losses  # noqa: F401  (makes yp.losses resolvable)
    done = []
    table = {k: dict(v) for k, v in BINDINGS.items()}
    if losses:
        for k, v in LOSS_BINDINGS.items():
            table.setdefault(k, {}).update(v)
    for mod_name, names in table.items():
        if modules is not None and mod_name not in modules:
            continue
        try:
            mod = sys.modules.get(mod_name) or importlib.import_module(mod_name)
        except Exception:
            continue
        for attr, ours in names.items():
            if hasattr(mod, attr):
                _SAVED.append((mod, attr, getattr(mod, attr)))
                setattr(mod, attr, _resolve(yp, ours))
                done.append(f"{mod_name}.{attr}")
    return done


def uninstall() -> None:
    while _SAVED:
        mod, attr, old = _SAVED.pop()
        setattr(mod, attr, old)
