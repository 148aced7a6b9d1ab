"""yolopoint_b200 -- the YOLOPoint hot path (UniBwTAS/YOLOPoint) on hand-written sm_100a kernels.

Public surface mirrors the reference's Python call signatures:

    Model, load_model                                   src/models/YOLOPoint.py, src/utils/utils.py:55-57
    non_max_suppression, xywh2xyxy                      src/utils/general_yolo.py
    flattenDetection, getPtsFromHeatmap, getPtsFromSemi, nms_fast     src/utils/utils.py
    sample_desc_from_points                             src/evaluations/descriptor_evaluation.py
    warp_image_batch, homography_adaptation             src/utils/utils.py:333-376, src/export_homography.py:97-128
    PointTracker (update / get_tracks / nn_match_two_way), nn_match_two_way     src/demo.py:268-441
    keypoints_to_wire, objects_to_wire                  src/yolopoint_ros.py:109-145 (KeypointArray.msg / ObjectInstance2D fields)
    YoloPointFrontend.process_img                       src/demo.py
    detect, extract_keypoints, match                    convenience names from BASELINE.json

Nothing here falls back to PyTorch or the CPU for inference: the CUDA library must be built and an sm_100
device present, otherwise the calls raise.
"""
from .model import Model, YOLOPoint, YOLOPointv52, load_model  # noqa: F401
from .api import (detect, extract_keypoints, flattenDetection, getPtsFromHeatmap, getPtsFromSemi, homography_adaptation, match,  # noqa: F401
                  nms_fast, nn_match_two_way, non_max_suppression, sample_desc_from_points, warp_image_batch, xywh2xyxy)
from .frontend import DEFAULT_CFG, FramePipeline, YoloPointFrontend  # noqa: F401
from .tracker import PointTracker, keypoints_from_wire, keypoints_to_wire, objects_to_wire  # noqa: F401

__version__ = "0.1.0"
