"""Import the UNMODIFIED reference (UniBwTAS/YOLOPoint) from /root/reference  --  TEST INFRASTRUCTURE.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference tree is absent).

Shims (none modify the reference):
  * stub ``matplotlib`` / ``matplotlib.pyplot`` (imported at module import by utils/metrics_yolo.py:10),
  * ``RANK=1`` so utils/plots_yolo.py:64-66 does not try to download a font,
  * no bytecode writes (the tree is read-only),
  * ``YoloPointFrontend`` is built with ``object.__new__`` because its ``__init__`` needs a checkpoint
    file (src/demo.py:34).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("YOLOPOINT_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "models"))


def load():
    """Returns a namespace with the reference symbols on the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    os.environ["RANK"] = "1"
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    sys.dont_write_bytecode = True
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot")
        mpl.use = lambda *a, **k: None; mpl.rc = lambda *a, **k: None; mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
    src = os.path.join(REF_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    ns = types.SimpleNamespace()
    from models.YOLOPoint import Model  # noqa
    from utils.general_yolo import non_max_suppression, xywh2xyxy  # noqa
    from utils.utils import flattenDetection, getPtsFromHeatmap, getPtsFromSemi, nms_fast  # noqa
    from evaluations.descriptor_evaluation import sample_desc_from_points  # noqa
    import demo  # noqa
    ns.Model, ns.non_max_suppression, ns.xywh2xyxy = Model, non_max_suppression, xywh2xyxy
    ns.flattenDetection, ns.getPtsFromHeatmap, ns.getPtsFromSemi, ns.nms_fast = \
        flattenDetection, getPtsFromHeatmap, getPtsFromSemi, nms_fast
    ns.sample_desc_from_points = sample_desc_from_points
    ns.PointTracker, ns.YoloPointFrontend = demo.PointTracker, demo.YoloPointFrontend
    return ns


def make_frontend(ns, model, cfg, filter_pts=True):
    """A reference YoloPointFrontend around ``model`` without touching a checkpoint (src/demo.py:18-49)."""
    fe = object.__new__(ns.YoloPointFrontend)
    fe.device = "cpu"
    fe.cell, fe.border_remove = 8, 4
    fe.sp_config = dict(detection_threshold=cfg["detection_threshold"], nms=cfg["nms"], nn_thresh=cfg["nn_thresh"])
    fe.yolo_config = dict(conf_thres_box=cfg["conf_thres_box"], iou_thres_box=cfg["iou_thres_box"], max_det=cfg["max_det"])
    fe.filter_pts, fe.crop_resize, fe.templates = filter_pts, None, {}
    fe.model = model
    return fe
