// Host build of yolopoint_b200/csrc/box_loss_math.cuh for the CPU unit test of the object-loss arithmetic
// (tests/test_losses.py::test_box_loss_math_matches_autograd).  Test infrastructure only: the product path is csrc/object_loss.cu.
#include "../yolopoint_b200/csrc/box_loss_math.cuh"

extern "C" void yp_host_candidate_ciou(const float* q, const float* anchor, const float* tbox, int n, float eps, float* ciou, float* grad) {
  for (int i = 0; i < n; ++i) ciou[i] = yp::candidate_ciou(q + 4 * i, anchor[2 * i], anchor[2 * i + 1], tbox + 4 * i, eps, grad + 4 * i);
}

extern "C" void yp_host_bce_logits(const float* x, const float* t, int n, float pw, float* loss, float* dx) {
  for (int i = 0; i < n; ++i) loss[i] = yp::bce_logits(x[i], t[i], pw, dx + i);
}
