"""Every convolution launch of every supported model / version / precision at the BASELINE.json input shapes passes the library's
host-side planner (yp_conv2d_plan_check: formats, channel and tile geometry, shared-memory and TMEM budgets) -- on the CPU, before
anything runs on a GPU.  Shapes: configs[0] N 640x480, configs[1] S 640x640, configs[2] M 1280x736 batch 4 per GPU, configs[4] L 640x640
batch 8 per GPU."""
import ctypes as C

import pytest

from yolopoint_b200 import _lib
from yolopoint_b200._lib import YP_ALGO_SIMT, YpConvDesc
from yolopoint_b200.engine import MODEL_NAMES, NetPlan, check_plan

SHAPES = {"n": (1, 480, 640), "s": (1, 640, 640), "m": (4, 736, 1280), "l": (8, 640, 640)}


@pytest.mark.parametrize("model_name", MODEL_NAMES)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("ver", sorted(SHAPES))
def test_every_launch_plans(model_name, precision, ver):
    net = NetPlan(ver, 80, precision, model_name)
    B, H, W = SHAPES[ver]
    assert check_plan(net, B, H, W) == []
    assert check_plan(net, 1, 64, 96) == []                  # the smallest frames the tests use


def test_planner_reports_what_is_wrong():
    net = NetPlan("s", 80, "fp32")
    op = net.conv_ops()[1]
    op.cout += 8                                            # not a multiple of 16
    bad = check_plan(net, 1, 64, 64)
    assert len(bad) >= 1 and bad[0][0] == "+".join(op.names) and "multiples of 16" in bad[0][1]
    L = _lib.lib()
    assert L.yp_conv2d_plan_check(None) != 0
    d = YpConvDesc()
    d.algo = YP_ALGO_SIMT
    assert L.yp_conv2d_plan_check(C.byref(d)) != 0 and b"tcgen05" in L.yp_last_error()


@pytest.mark.parametrize("model_name", MODEL_NAMES)
def test_version_x_plans_with_a_separate_l2norm(model_name):
    """D = 320 exceeds the single-tile L2-norm epilogue: the last descriptor convolution is planned without it, followed by the
    in-place row normalisation (yp_l2norm_nhwc); every launch of the plan passes the library's dry run."""
    from yolopoint_b200.engine import L2NormOp
    net = NetPlan("x", 80, "fp32", model_name)
    assert net.D == 320 and not net.fused_l2norm
    assert sum(isinstance(op, L2NormOp) for op in net.ops) == 1 and not any(op.l2norm for op in net.conv_ops())
    assert check_plan(net, 1, 640, 640) == []
    assert NetPlan("l", 80, "fp32", model_name).fused_l2norm
