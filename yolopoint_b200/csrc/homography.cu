// Homography-adaptation export: inverse-homography warp of a batch of images with bilinear / nearest
// sampling (utils/utils.py:333-376 `warp_image_batch` = warp_points :274-290 + F.grid_sample(align_corners=True, zeros padding))
// and the aggregation of export_homography.py:97-128:
//     heat_b * mask_b and mask_b are warped back with the inverse homography of copy b, summed over the B copies and divided.
// The reference materialises four [B,1,H,W] tensors (product, two warps, coordinates) and reduces them afterwards; here one thread
// owns one output pixel and walks the B copies: heat and mask are each read once (4 taps per copy, neighbouring threads read
// neighbouring source pixels), nothing but the [H,W] result is written.  HBM-bound: 2 * B * H * W * 4 bytes in, H * W * 4 out.
#include "common.cuh"

namespace yp {
namespace {

struct Src { float ix, iy; };

// source pixel of output pixel (xn, yn) (normalised linspace coordinates) under the inverse homography h[9]
__device__ __forceinline__ Src source_coords(const float* __restrict__ h, float xn, float yn, int H, int W) {
  const float p0 = fmaf(h[2], 1.0f, fmaf(h[1], yn, __fmul_rn(h[0], xn)));
  const float p1 = fmaf(h[5], 1.0f, fmaf(h[4], yn, __fmul_rn(h[3], xn)));
  const float p2 = fmaf(h[8], 1.0f, fmaf(h[7], yn, __fmul_rn(h[6], xn)));
  const float u = __fdiv_rn(p0, p2), v = __fdiv_rn(p1, p2);
  Src s;   // grid_sampler_unnormalize, align_corners = True: ((g + 1) / 2) * (size - 1)
  s.ix = __fmul_rn(__fdiv_rn(__fadd_rn(u, 1.0f), 2.0f), static_cast<float>(W - 1));
  s.iy = __fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), static_cast<float>(H - 1));
  return s;
}

struct Taps {
  int off[4];      // element offsets of nw, ne, sw, se (clamped to a valid address)
  float w[4];      // bilinear weights, 0 for out-of-range corners (zeros padding)
};

__device__ __forceinline__ Taps bilinear_taps(Src s, int H, int W) {
  Taps t;
  const float x0 = floorf(s.ix), y0 = floorf(s.iy);
  const float x1 = __fadd_rn(x0, 1.0f), y1 = __fadd_rn(y0, 1.0f);
  const float wx0 = __fsub_rn(x1, s.ix), wx1 = __fsub_rn(s.ix, x0), wy0 = __fsub_rn(y1, s.iy), wy1 = __fsub_rn(s.iy, y0);
  const float xs[2] = {x0, x1}, ys[2] = {y0, y1};
  const float wxs[2] = {wx0, wx1}, wys[2] = {wy0, wy1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float xq = xs[k & 1], yq = ys[k >> 1];
    const bool ok = xq >= 0.0f && xq <= static_cast<float>(W - 1) && yq >= 0.0f && yq <= static_cast<float>(H - 1);   // false for NaN / inf
    const int xi = ok ? static_cast<int>(xq) : 0, yi = ok ? static_cast<int>(yq) : 0;
    t.off[k] = yi * W + xi;
    t.w[k] = ok ? __fmul_rn(wxs[k & 1], wys[k >> 1]) : 0.0f;
  }
  return t;
}

// ---- warp_image_batch: img [B,C,H,W] -> out [B,C,H,W]; Hinv [B,3,3]; xs[W], ys[H] = torch.linspace(-1, 1, n)
__global__ void __launch_bounds__(256) warp_batch_kernel(const float* __restrict__ img, const float* __restrict__ Hinv, const float* __restrict__ xs,
                                                         const float* __restrict__ ys, int B, int Cc, int H, int W, int nearest,
                                                         float* __restrict__ out) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t HW = static_cast<int64_t>(H) * W;
  if (idx >= B * HW) return;
  const int b = static_cast<int>(idx / HW);
  const int pix = static_cast<int>(idx - b * HW);
  const int y = pix / W, x = pix - y * W;
  const Src s = source_coords(Hinv + b * 9, xs[x], ys[y], H, W);
  const float* src = img + static_cast<int64_t>(b) * Cc * HW;
  float* dst = out + static_cast<int64_t>(b) * Cc * HW + pix;
  if (nearest) {   // nearbyint = round half to even
    const float xr = rintf(s.ix), yr = rintf(s.iy);
    const bool ok = xr >= 0.0f && xr <= static_cast<float>(W - 1) && yr >= 0.0f && yr <= static_cast<float>(H - 1);
    const int off = ok ? static_cast<int>(yr) * W + static_cast<int>(xr) : 0;
    for (int c = 0; c < Cc; ++c) dst[c * HW] = ok ? __ldg(src + c * HW + off) : 0.0f;
    return;
  }
  const Taps t = bilinear_taps(s, H, W);
  for (int c = 0; c < Cc; ++c) {
    const float* sc = src + c * HW;
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc = __fadd_rn(acc, __fmul_rn(__ldg(sc + t.off[k]), t.w[k]));
    dst[c * HW] = acc;
  }
}

// ---- fused aggregation: heat, mask [B,H,W] -> agg [H,W] = sum_b warp(heat_b * mask_b) / sum_b warp(mask_b)
constexpr int kMaxCopies = 1024;

__global__ void __launch_bounds__(128) homography_adapt_kernel(const float* __restrict__ heat, const float* __restrict__ mask,
                                                               const float* __restrict__ Hinv, const float* __restrict__ xs,
                                                               const float* __restrict__ ys, int B, int H, int W, float* __restrict__ sum_h,
                                                               float* __restrict__ sum_m, float* __restrict__ agg) {
  extern __shared__ float sh[];   // [B][9]
  for (int i = threadIdx.x; i < B * 9; i += blockDim.x) sh[i] = Hinv[i];
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (pix >= HW) return;
  const int y = pix / W, x = pix - y * W;
  const float xn = xs[x], yn = ys[y];
  float ah = 0.0f, am = 0.0f;
  for (int b = 0; b < B; ++b) {
    const Taps t = bilinear_taps(source_coords(sh + b * 9, xn, yn, H, W), H, W);
    const float* hb = heat + static_cast<int64_t>(b) * HW;
    const float* mb = mask + static_cast<int64_t>(b) * HW;
    float vh[4], vm[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { vh[k] = __ldg(hb + t.off[k]); vm[k] = __ldg(mb + t.off[k]); }
    float wh = 0.0f, wm = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      wh = __fadd_rn(wh, __fmul_rn(__fmul_rn(vh[k], vm[k]), t.w[k]));   // heatmap * mask_2D, then the warp (export_homography.py:98-99)
      wm = __fadd_rn(wm, __fmul_rn(vm[k], t.w[k]));
    }
    ah = __fadd_rn(ah, wh);
    am = __fadd_rn(am, wm);
  }
  if (sum_h) sum_h[pix] = ah;
  if (sum_m) sum_m[pix] = am;
  agg[pix] = __fdiv_rn(ah, am);   // 0 / 0 = NaN where no copy covers the pixel, as in the reference
}

}  // namespace
}  // namespace yp

extern "C" int yp_warp_image_batch(const float* img, const float* hinv, const float* xs, const float* ys, int32_t B, int32_t C, int32_t H, int32_t W,
                                   int32_t nearest, float* out, void* stream) {
  YP_REQUIRE(img && hinv && xs && ys && out, YP_ERR_ARG, "warp_image_batch: null pointer");
  YP_REQUIRE(B > 0 && C > 0 && H > 1 && W > 1 && static_cast<int64_t>(H) * W < (1ll << 31), YP_ERR_SHAPE, "warp_image_batch: B=%d C=%d H=%d W=%d", B, C, H, W);
  const int64_t n = static_cast<int64_t>(B) * H * W;
  yp::warp_batch_kernel<<<static_cast<unsigned>(yp::ceil_div64(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, hinv, xs, ys, B, C, H, W,
                                                                                                               nearest, out);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_homography_adaptation(const float* heat, const float* mask, const float* hinv, const float* xs, const float* ys, int32_t B,
                                        int32_t H, int32_t W, float* sum_heat, float* sum_mask, float* agg, void* stream) {
  YP_REQUIRE(heat && mask && hinv && xs && ys && agg, YP_ERR_ARG, "homography_adaptation: null pointer");
  YP_REQUIRE(B > 0 && B <= yp::kMaxCopies && H > 1 && W > 1 && static_cast<int64_t>(H) * W < (1ll << 31), YP_ERR_SHAPE,
             "homography_adaptation: B=%d (<= %d) H=%d W=%d", B, yp::kMaxCopies, H, W);
  yp::homography_adapt_kernel<<<yp::ceil_div(H * W, 128), 128, sizeof(float) * 9 * B, static_cast<cudaStream_t>(stream)>>>(heat, mask, hinv, xs, ys, B, H,
                                                                                                                       W, sum_heat, sum_mask, agg);
  YP_LAUNCH_OK();
  return YP_OK;
}
