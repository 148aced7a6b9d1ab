// Host build of yolopoint_b200/csrc/box_loss_math.cuh for the CPU unit test of the object-loss arithmetic
// (tests/test_losses.py::test_box_loss_math_matches_autograd).  Test infrastructure only: the product path is csrc/object_loss.cu.
#include "../yolopoint_b200/csrc/box_loss_math.cuh"

extern "C" void yp_host_candidate_ciou(const float* q, const float* anchor, const float* tbox, int n, float eps, float* ciou, float* grad) {
  for (int i = 0; i < n; ++i) ciou[i] = yp::candidate_ciou(q + 4 * i, anchor[2 * i], anchor[2 * i + 1], tbox + 4 * i, eps, grad + 4 * i);
}

extern "C" void yp_host_bce_logits(const float* x, const float* t, int n, float pw, float* loss, float* dx) {
  for (int i = 0; i < n; ++i) loss[i] = yp::bce_logits(x[i], t[i], pw, dx + i);
}

// plan of all 5 x na x nt candidates of one level in the order of build_targets: e = (o * na + a) * nt + t
extern "C" void yp_host_plan_candidates(const float* targets, int nt, const float* anchors, int na, int nx, int ny, float anchor_t, unsigned char* valid,
                                        long long* cell, float* tbox, int* cls) {
  for (int o = 0; o < 5; ++o)
    for (int a = 0; a < na; ++a)
      for (int t = 0; t < nt; ++t) {
        const long long e = (static_cast<long long>(o) * na + a) * nt + t;
        const yp::CandPlan c = yp::plan_candidate(targets + 6 * t, anchors[2 * a], anchors[2 * a + 1], nx, ny, o, anchor_t);
        valid[e] = c.valid;
        cell[e] = ((static_cast<long long>(c.img) * na + a) * ny + c.gj) * nx + c.gi;
        for (int k = 0; k < 4; ++k) tbox[4 * e + k] = c.tbox[k];
        cls[e] = c.cls;
      }
}
