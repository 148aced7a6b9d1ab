// Per-candidate arithmetic of the object loss (YOLOv5 box / class terms as the reference configures them,
// src/utils/loss_functions.py:174-215 with bbox_iou(..., CIoU=True) of src/utils/metrics_yolo.py:202-240): value AND gradient with
// respect to the four raw box logits in one evaluation.  The gradient is carried forward through the CIoU expression by a small
// dual-number type (value + 4 partial derivatives), so there is no hand-derived formula to get wrong; `alpha` is a constant of the
// backward pass as in the reference (computed under no_grad there).
//
// Plain C++ under YP_HD so that the same lines compile for the device (csrc/object_loss.cu) and, for the CPU unit test of the
// arithmetic, for the host (tests/box_loss_host.cpp against PyTorch autograd) -- the host build is test infrastructure only.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define YP_HD __host__ __device__ __forceinline__
#else
#define YP_HD inline
#endif

namespace yp {

struct Dual4 {
  float v;
  float d[4];
};

YP_HD Dual4 dconst(float v) { return Dual4{v, {0.f, 0.f, 0.f, 0.f}}; }
YP_HD Dual4 dvar(float v, int i) {
  Dual4 r = dconst(v);
  r.d[i] = 1.f;
  return r;
}
YP_HD Dual4 operator+(const Dual4& a, const Dual4& b) { return Dual4{a.v + b.v, {a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2], a.d[3] + b.d[3]}}; }
YP_HD Dual4 operator-(const Dual4& a, const Dual4& b) { return Dual4{a.v - b.v, {a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2], a.d[3] - b.d[3]}}; }
YP_HD Dual4 operator*(const Dual4& a, const Dual4& b) {
  return Dual4{a.v * b.v, {a.d[0] * b.v + a.v * b.d[0], a.d[1] * b.v + a.v * b.d[1], a.d[2] * b.v + a.v * b.d[2], a.d[3] * b.v + a.v * b.d[3]}};
}
YP_HD Dual4 operator/(const Dual4& a, const Dual4& b) {
  const float inv = 1.f / b.v, q = a.v * inv;
  return Dual4{q, {(a.d[0] - q * b.d[0]) * inv, (a.d[1] - q * b.d[1]) * inv, (a.d[2] - q * b.d[2]) * inv, (a.d[3] - q * b.d[3]) * inv}};
}
YP_HD Dual4 operator+(const Dual4& a, float b) { Dual4 r = a; r.v += b; return r; }
YP_HD Dual4 operator-(const Dual4& a, float b) { Dual4 r = a; r.v -= b; return r; }
YP_HD Dual4 operator-(float a, const Dual4& b) { return Dual4{a - b.v, {-b.d[0], -b.d[1], -b.d[2], -b.d[3]}}; }
YP_HD Dual4 operator*(const Dual4& a, float b) { return Dual4{a.v * b, {a.d[0] * b, a.d[1] * b, a.d[2] * b, a.d[3] * b}}; }
YP_HD Dual4 dmin(const Dual4& a, float b) { return a.v <= b ? a : dconst(b); }
YP_HD Dual4 dmax(const Dual4& a, float b) { return a.v >= b ? a : dconst(b); }
YP_HD Dual4 dclamp0(const Dual4& a) { return a.v >= 0.f ? a : dconst(0.f); }
YP_HD Dual4 datan(const Dual4& a) {
  const float g = 1.f / (1.f + a.v * a.v);
  return Dual4{atanf(a.v), {a.d[0] * g, a.d[1] * g, a.d[2] * g, a.d[3] * g}};
}

YP_HD float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Complete-IoU of the predicted box (px, py, pw, ph as duals over the four raw logits) against the target box t (constants),
// both in centre format, with eps where the reference adds it (union, c^2, the two aspect ratios, alpha).
YP_HD Dual4 ciou_dual(const Dual4& px, const Dual4& py, const Dual4& pw, const Dual4& ph, float tx, float ty, float tw, float th, float eps) {
  const Dual4 p_l = px - pw * 0.5f, p_r = px + pw * 0.5f, p_t = py - ph * 0.5f, p_b = py + ph * 0.5f;
  const float t_l = tx - tw * 0.5f, t_r = tx + tw * 0.5f, t_t = ty - th * 0.5f, t_b = ty + th * 0.5f;
  const Dual4 inter = dclamp0(dmin(p_r, t_r) - dmax(p_l, t_l)) * dclamp0(dmin(p_b, t_b) - dmax(p_t, t_t));
  const Dual4 iou = inter / ((pw * ph + tw * th) - inter + eps);
  const Dual4 cw = dmax(p_r, t_r) - dmin(p_l, t_l), chh = dmax(p_b, t_b) - dmin(p_t, t_t);
  const Dual4 diag2 = cw * cw + chh * chh + eps;
  const Dual4 ex = (t_l + t_r) - (p_l + p_r), ey = (t_t + t_b) - (p_t + p_b);
  const Dual4 rho2 = (ex * ex + ey * ey) * 0.25f;
  const Dual4 da = atanf(tw / (th + eps)) - datan(pw / (ph + eps));
  const Dual4 v = (da * da) * 0.40528473456935109f;            // 4 / pi^2
  const float alpha = v.v / (v.v - iou.v + (1.f + eps));        // constant of the backward pass
  return iou - (rho2 / diag2 + v * alpha);
}

// One assignment candidate: raw box logits q[0..3], anchor (w, h) in cells, target box (x, y relative to the cell, w, h in cells).
// -> CIoU; grad[k] = d CIoU / d q[k].  Box decoding as src/utils/loss_functions.py:176-178: xy = 2 sigma(q) - 0.5, wh = (2 sigma(q))^2 * anchor.
YP_HD float candidate_ciou(const float* q, float aw, float ah, const float* tbox, float eps, float* grad) {
  const float s0 = sigmoidf_(q[0]), s1 = sigmoidf_(q[1]), s2 = sigmoidf_(q[2]), s3 = sigmoidf_(q[3]);
  // d xy / d q = 2 s (1 - s);  d wh / d q = 8 s^2 (1 - s) anchor: seeded straight into the duals
  Dual4 px = dconst(2.f * s0 - 0.5f), py = dconst(2.f * s1 - 0.5f), pw = dconst(4.f * s2 * s2 * aw), ph = dconst(4.f * s3 * s3 * ah);
  px.d[0] = 2.f * s0 * (1.f - s0);
  py.d[1] = 2.f * s1 * (1.f - s1);
  pw.d[2] = 8.f * s2 * s2 * (1.f - s2) * aw;
  ph.d[3] = 8.f * s3 * s3 * (1.f - s3) * ah;
  const Dual4 c = ciou_dual(px, py, pw, ph, tbox[0], tbox[1], tbox[2], tbox[3], eps);
  grad[0] = c.d[0]; grad[1] = c.d[1]; grad[2] = c.d[2]; grad[3] = c.d[3];
  return c.v;
}

// Target assignment of one candidate (ComputeObjectLoss.build_targets, src/utils/loss_functions.py:218-234 in its fixed-shape form):
// candidate = (offset variant o: centre, left, up, right, down; anchor (aw, ah) in cells; target tg = (image, class, x, y, w, h) with
// the box normalised to [0, 1]) on an nx x ny grid.  Valid when no side of the target is more than anchor_t times longer or shorter
// than the anchor's and, for o > 0, when the centre lies in the matching half of its cell and more than one cell away from that
// border.  Every operation is the single fp32 operation the PyTorch statement performs (multiply, divide, remainder, truncation),
// so borderline decisions agree bit for bit.
struct CandPlan {
  bool valid;
  int img, cls, gi, gj;     // gi, gj clamped to the grid
  float tbox[4];            // (x, y) relative to the UNclamped cell, (w, h) in cells
};

YP_HD float mul_rn(float a, float b) {   // a product that is rounded on its own (never contracted into an FMA with a following add)
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}

YP_HD float remainder1(float x) {   // torch.remainder(x, 1): the sign follows the divisor
  float r = fmodf(x, 1.f);
  if (r != 0.f && r < 0.f) r += 1.f;
  return r;
}

YP_HD CandPlan plan_candidate(const float* tg, float aw, float ah, int nx, int ny, int o, float anchor_t) {
  CandPlan c;
  const float fx = static_cast<float>(nx), fy = static_cast<float>(ny);
  const float gx = mul_rn(tg[2], fx), gy = mul_rn(tg[3], fy), gw = mul_rn(tg[4], fx), gh = mul_rn(tg[5], fy);
  const float rw = gw / aw, rh = gh / ah;
  const float worst = fmaxf(fmaxf(rw, 1.f / rw), fmaxf(rh, 1.f / rh));
  bool ok = worst < anchor_t;
  float hx = 0.f, hy = 0.f;
  if (o == 1) { ok = ok && remainder1(gx) < 0.5f && gx > 1.f; hx = 0.5f; }
  else if (o == 2) { ok = ok && remainder1(gy) < 0.5f && gy > 1.f; hy = 0.5f; }
  else if (o == 3) { const float ix = fx - gx; ok = ok && remainder1(ix) < 0.5f && ix > 1.f; hx = -0.5f; }
  else if (o == 4) { const float iy = fy - gy; ok = ok && remainder1(iy) < 0.5f && iy > 1.f; hy = -0.5f; }
  const long long cx = static_cast<long long>(gx - hx), cy = static_cast<long long>(gy - hy);   // truncation toward zero, like .long()
  c.valid = ok;
  c.img = static_cast<int>(static_cast<long long>(tg[0]));
  c.cls = static_cast<int>(static_cast<long long>(tg[1]));
  c.gi = static_cast<int>(cx < 0 ? 0 : cx > nx - 1 ? nx - 1 : cx);
  c.gj = static_cast<int>(cy < 0 ? 0 : cy > ny - 1 ? ny - 1 : cy);
  c.tbox[0] = gx - static_cast<float>(cx);
  c.tbox[1] = gy - static_cast<float>(cy);
  c.tbox[2] = gw;
  c.tbox[3] = gh;
  return c;
}

// BCE-with-logits term with a positive-class weight (torch.nn.functional.binary_cross_entropy_with_logits(pos_weight=pw)):
// L = (1 - t) x + (1 + (pw - 1) t) softplus(-x);  *dx = dL/dx.
YP_HD float bce_logits(float x, float t, float pw, float* dx) {
  const float w = 1.f + (pw - 1.f) * t;
  const float sp = log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f);
  *dx = (1.f - t) - w * (1.f - sigmoidf_(x));
  return (1.f - t) * x + w * sp;
}

}  // namespace yp
