"""ctypes binding of libyolopoint_b200.so (the C ABI declared in include/yolopoint_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  ``lib()`` raises if the shared object
is missing, if a symbol declared in the header is not exported, or (``require_device``) if the current
device is not an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libyolopoint_b200.so")

YP_FMT_F32X2, YP_FMT_BF16, YP_FMT_F32 = 0, 1, 2
YP_ACT_NONE, YP_ACT_SILU = 0, 1
YP_ALGO_TCGEN05, YP_ALGO_SIMT = 0, 1
YP_EPI_L2NORM = 1
YP_EPI_ROWMIN = 4
YP_UP_PARITY = 16
YP_CAT_COPY, YP_CAT_UP2, YP_CAT_POOL2 = 0, 1, 2
STATUS = {0: "YP_OK", -1: "YP_ERR_SHAPE", -2: "YP_ERR_ALIGN", -3: "YP_ERR_ARCH", -4: "YP_ERR_CUDA",
          -5: "YP_ERR_CAPACITY", -6: "YP_ERR_ARG"}


class YpView(C.Structure):
    _fields_ = [("base", C.c_void_p), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("pix_stride", C.c_int64), ("plane_stride", C.c_int64), ("format", C.c_int32), ("upsample", C.c_int32)]


class YpConvDesc(C.Structure):
    _fields_ = [("in_", YpView), ("weight", C.c_void_p), ("bias", C.c_void_p), ("ksize", C.c_int32), ("stride", C.c_int32),
                ("cout", C.c_int32), ("act", C.c_int32), ("epilogue", C.c_uint32), ("residual", YpView), ("n_out", C.c_int32),
                ("out", YpView * 2), ("algo", C.c_int32), ("tile_n", C.c_int32), ("split_k", C.c_int32), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_uint64), ("n_taps", C.c_int32), ("tap_dh", C.c_int8 * 9), ("tap_dw", C.c_int8 * 9),
                ("row_key", C.c_void_p), ("n_rows", C.c_void_p), ("n_cols", C.c_void_p), ("col_off", C.c_int32), ("col_key", C.c_void_p)]


class YpChainOp(C.Structure):
    _fields_ = [("type", C.c_int32), ("conv", YpConvDesc), ("pool", YpView), ("n_deps", C.c_int32), ("deps", C.c_int32 * 8)]


class YpWgradDesc(C.Structure):
    _fields_ = [("x", YpView), ("dy", YpView), ("ksize", C.c_int32), ("stride", C.c_int32), ("dw", C.c_void_p)]


class YpNmsParams(C.Structure):
    _fields_ = [("conf_thres", C.c_float), ("iou_thres", C.c_float), ("multi_label", C.c_int32), ("agnostic", C.c_int32),
                ("max_det", C.c_int32), ("max_nms", C.c_int32), ("max_wh", C.c_float), ("class_mask", C.c_void_p)]


class YpCatPart(C.Structure):
    _fields_ = [("src", C.c_void_p), ("grad", C.c_void_p), ("C", C.c_int32), ("mode", C.c_int32)]


class YpObjLossLevel(C.Structure):
    _fields_ = [("pred", C.c_void_p), ("dpred", C.c_void_p), ("valid", C.c_void_p), ("cell", C.c_void_p), ("tbox", C.c_void_p),
                ("anchor", C.c_void_p), ("cls", C.c_void_p), ("cells", C.c_int64), ("E", C.c_int32), ("balance", C.c_float),
                ("targets", C.c_void_p), ("nt", C.c_int32), ("na", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("nb", C.c_int32),
                ("anchors", C.c_float * 16)]


class YpObjLossParams(C.Structure):
    _fields_ = [("cp", C.c_float), ("cn", C.c_float), ("cls_pw", C.c_float), ("obj_pw", C.c_float), ("gr", C.c_float),
                ("w_box", C.c_float), ("w_obj", C.c_float), ("w_cls", C.c_float), ("eps", C.c_float), ("anchor_t", C.c_float)]


_i32, _i64, _f32, _vp, _sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t
_PV, _PC, _PN = C.POINTER(YpView), C.POINTER(YpConvDesc), C.POINTER(YpNmsParams)

# name -> (restype, argtypes); must list every symbol of include/yolopoint_b200.h
SIGNATURES = {
    "yp_abi_version": (_i32, []),
    "yp_last_error": (C.c_char_p, []),
    "yp_check_device": (_i32, []),
    "yp_memcpy_async": (_i32, [_vp, _vp, _sz, _vp]),
    "yp_conv2d_nhwc_fwd": (_i32, [_PC, _vp]),
    "yp_conv2d_workspace_bytes": (_sz, [_PC]),
    "yp_conv2d_plan_check": (_i32, [_PC]),
    "yp_conv_chain_create": (_i32, [C.POINTER(YpChainOp), _i32, C.POINTER(C.c_void_p)]),
    "yp_conv_chain_launch": (_i32, [_vp, _vp]),
    "yp_conv_chain_destroy": (_i32, [_vp]),
    "yp_conv_chain_info": (_i32, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "yp_debug_conv_chain_timeline": (_i32, [_vp, _i32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "yp_conv2d_nhwc_wgrad": (_i32, [C.POINTER(YpWgradDesc), _vp]),
    "yp_bn_act_fwd": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _f32, _f32, _i32, _vp, _vp, _vp, _vp]),
    "yp_bn_act_bwd": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _vp]),
    "yp_debug_conv_timeline": (_i32, [_vp]),
    "yp_sppf_pool": (_i32, [_PV, _vp]),
    "yp_l2norm_nhwc": (_i32, [_PV, _vp]),
    "yp_split_tf32": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "yp_maxpool2x2": (_i32, [_PV, _PV, _vp]),
    "yp_nchw_to_s2d": (_i32, [_vp, _i32, _i32, _i32, _PV, _vp]),
    "yp_frame_to_s2d": (_i32, [_vp, _i32, _i32, _i32, _PV, _vp]),
    "yp_nhwc_to_nchw": (_i32, [_PV, _i32, _vp, _vp]),
    "yp_detect_decode": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, C.POINTER(C.c_float), _vp, _vp, _i64, _i64, _vp]),
    "yp_box_nms_workspace_bytes": (_sz, [_i32, _i64, _i32, _i32]),
    "yp_box_nms": (_i32, [_vp, _i32, _i64, _i32, _PN, _i32, _vp, _vp, _vp, _sz, _vp]),
    "yp_detect_nms": (_i32, [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float),
                             C.POINTER(C.c_float), _i32, _i32, _i32, _PN, _i32, _vp, _vp, _vp, _sz, _i32, _vp]),
    "yp_detect_prescan": (_i32, [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), _i32, _i32, _i32, _PN, _i32, _i32, _vp, _sz, _vp]),
    "yp_heatmap": (_i32, [_vp, _i32, _i32, _i32, _i64, _i64, _i64, _i64, _i32, _vp, _vp]),
    "yp_keypoints_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "yp_keypoints_nms": (_i32, [_vp, _i32, _i32, _i32, _f32, _i32, _i32, _vp, _sz, _vp]),
    "yp_keypoints_collect": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _sz, _vp]),
    "yp_keypoints_filter": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "yp_keypoints_threshold_count": (_i32, [_vp, _sz, _i32, _i32, _i32, _i32, _vp, _vp]),
    "yp_match_frames": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "yp_gather_rows": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "yp_keypoints": (_i32, [_vp, _i32, _i32, _i32, _f32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _sz, _vp]),
    "yp_sample_desc": (_i32, [_vp, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "yp_warp_image_batch": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "yp_homography_adaptation": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "yp_detector_loss_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "yp_detector_loss": (_i32, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "yp_object_loss_workspace_bytes": (_sz, [C.POINTER(YpObjLossLevel), _i32]),
    "yp_object_loss": (_i32, [C.POINTER(YpObjLossLevel), _i32, _i32, _i32, C.POINTER(YpObjLossParams), _vp, _vp, _sz, _vp]),
    "yp_cat_nhwc_fwd": (_i32, [C.POINTER(YpCatPart), _i32, _vp, _i32, _i32, _i32, _vp]),
    "yp_cat_nhwc_bwd": (_i32, [C.POINTER(YpCatPart), _i32, _vp, _i32, _i32, _i32, _vp]),
    "yp_sppf_train_fwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "yp_sppf_train_bwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "yp_match_partial": (_i32, [_vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "yp_match_finalize": (_i32, [_vp, _vp, _i32, _vp, _i32, _f32, _vp, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None
_device_ok = set()


class YoloPointB200Error(RuntimeError):
    pass


def lib(require_device: bool = False):
    """Load (once) and return the shared library; fail loudly when it is unusable."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise YoloPointB200Error(
                        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        f"(or `make -C yolopoint_b200/csrc`). yolopoint_b200 has no CPU/PyTorch fallback.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    try:
                        fn = getattr(handle, name)
                    except AttributeError as e:  # pragma: no cover
                        raise YoloPointB200Error(f"{LIB_PATH} does not export {name}") from e
                    fn.restype, fn.argtypes = res, args
                if handle.yp_abi_version() != 9:
                    raise YoloPointB200Error("ABI version mismatch between _lib.py and libyolopoint_b200.so")
                _lib = handle
    if require_device:
        import torch
        if not torch.cuda.is_available():
            raise YoloPointB200Error("yolopoint_b200 needs a CUDA device (B200, sm_100a); none is visible")
        dev = torch.cuda.current_device()
        if dev not in _device_ok:
            check(_lib.yp_check_device())
            _device_ok.add(dev)
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = _lib.yp_last_error().decode("utf-8", "replace") if _lib is not None else ""
        if rc == -6 and ("Invalid" in msg or "nn_thresh" in msg):
            # the reference raises AssertionError / ValueError for these (general_yolo.py:149-150, demo.py:317-318)
            raise (ValueError(msg) if "nn_thresh" in msg else AssertionError(msg))
        raise YoloPointB200Error(f"{STATUS.get(rc, rc)}: {msg}")
