// Layout / pooling kernels: input conversion to the stem's space-to-depth NHWC operand, NHWC -> NCHW
// export of network outputs, and the single-pass SPPF pooling.  All are HBM/L2-bound element movers.
#include "common.cuh"

namespace yp {
namespace {

// out[b, h2, w2, (ph*2+pw)*3 + c] = src(b, c, 2*h2+ph, 2*w2+pw); channels 12..15 = 0.
template <bool kFrame>
__global__ void to_s2d_kernel(const void* __restrict__ src, int B, int H, int W, YpView out) {
  const int H2 = H / 2, W2 = W / 2;
  const int64_t total = static_cast<int64_t>(B) * H2 * W2;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int w2 = static_cast<int>(idx % W2);
  const int h2 = static_cast<int>((idx / W2) % H2);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(W2) * H2));
  float v[16];
#pragma unroll
  for (int i = 12; i < 16; ++i) v[i] = 0.0f;
#pragma unroll
  for (int ph = 0; ph < 2; ++ph)
#pragma unroll
    for (int pw = 0; pw < 2; ++pw)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h = 2 * h2 + ph, w = 2 * w2 + pw;
        float x;
        if (kFrame) {
          const uint8_t* f = static_cast<const uint8_t*>(src);
          x = __fdiv_rn(static_cast<float>(f[((static_cast<int64_t>(b) * H + h) * W + w) * 3 + c]), 255.0f);
        } else {
          const float* f = static_cast<const float*>(src);
          x = f[((static_cast<int64_t>(b) * 3 + c) * H + h) * W + w];
        }
        v[(ph * 2 + pw) * 3 + c] = x;
      }
  const int64_t off = idx * out.pix_stride;
  if (out.format == YP_FMT_BF16) {
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out.base) + off;
    uint4 pk[2];
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
    for (int i = 0; i < 8; ++i) p2[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    reinterpret_cast<uint4*>(o)[0] = pk[0];
    reinterpret_cast<uint4*>(o)[1] = pk[1];
  } else {
    float* o = static_cast<float*>(out.base) + off;
    float hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      hi[i] = out.format == YP_FMT_F32X2 ? tf32_round(v[i]) : v[i];
      lo[i] = tf32_round(v[i] - hi[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(o)[i] = make_float4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
    if (out.format == YP_FMT_F32X2) {
      float* ol = o + out.plane_stride;
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(ol)[i] = make_float4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
    }
  }
}

int to_s2d(const void* src, bool frame, int B, int H, int W, const YpView* out, cudaStream_t st) {
  YP_REQUIRE(src && out && out->base, YP_ERR_ARG, "to_s2d: null pointer");
  YP_REQUIRE(H % 2 == 0 && W % 2 == 0 && B > 0, YP_ERR_SHAPE, "to_s2d: H=%d W=%d must be even", H, W);
  YP_REQUIRE(out->B == B && out->H == H / 2 && out->W == W / 2 && out->C == 16, YP_ERR_SHAPE,
             "to_s2d: output view must be [%d,%d,%d,16]", B, H / 2, W / 2);
  YP_REQUIRE(aligned16(out->base) && out->pix_stride % 8 == 0 && out->plane_stride % 4 == 0, YP_ERR_ALIGN, "to_s2d: output not 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * (H / 2) * (W / 2);
  const unsigned blocks = static_cast<unsigned>(ceil_div64(total, 256));
  if (frame) to_s2d_kernel<true><<<blocks, 256, 0, st>>>(src, B, H, W, *out);
  else to_s2d_kernel<false><<<blocks, 256, 0, st>>>(src, B, H, W, *out);
  YP_LAUNCH_OK();
  return YP_OK;
}

// NHWC view -> dense NCHW fp32 via a 32(pixel) x 32(channel) shared-memory transpose.
__global__ void nhwc_to_nchw_kernel(YpView in, int C, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int64_t HW = static_cast<int64_t>(in.H) * in.W;
  const int b = blockIdx.z;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t p = p0 + i;
    const int c = c0 + threadIdx.x;
    float v = 0.0f;
    if (p < HW && c < C) v = load_act(in.base, in.format, in.plane_stride, (b * HW + p) * in.pix_stride + c);
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t p = p0 + threadIdx.x;
    if (p < HW && c < C) out[(static_cast<int64_t>(b) * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// SPPF: one pass over the 13x13 neighbourhood produces the 5x5 / 9x9 / 13x13 clipped-window maxima, which
// equal the three chained MaxPool2d(5,1,2) outputs (padding is -inf, so clipping commutes with chaining).
__global__ void sppf_pool_kernel(YpView cat4, int C) {
  const int64_t total = static_cast<int64_t>(cat4.B) * cat4.H * cat4.W * C;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  int64_t p = idx / C;
  const int w = static_cast<int>(p % cat4.W); p /= cat4.W;
  const int h = static_cast<int>(p % cat4.H);
  const int b = static_cast<int>(p / cat4.H);
  float best[3] = {-INFINITY, -INFINITY, -INFINITY};
  int64_t arg[3] = {0, 0, 0};
  for (int dy = -6; dy <= 6; ++dy) {
    const int y = h + dy;
    if (y < 0 || y >= cat4.H) continue;
    for (int dx = -6; dx <= 6; ++dx) {
      const int x = w + dx;
      if (x < 0 || x >= cat4.W) continue;
      const int64_t off = ((static_cast<int64_t>(b) * cat4.H + y) * cat4.W + x) * cat4.pix_stride + c;
      const float v = load_act(cat4.base, cat4.format, cat4.plane_stride, off);
      const int ring = max(abs(dy), abs(dx));  // <=2 -> all three windows, <=4 -> 9x9 and 13x13, else 13x13
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (ring <= 2 * (k + 1) && v > best[k]) { best[k] = v; arg[k] = off; }
    }
  }
  const int64_t self = ((static_cast<int64_t>(b) * cat4.H + h) * cat4.W + w) * cat4.pix_stride + c;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int64_t dst = self + static_cast<int64_t>(k + 1) * C;
    if (cat4.format == YP_FMT_BF16) {
      __nv_bfloat16* f = static_cast<__nv_bfloat16*>(cat4.base);
      f[dst] = f[arg[k]];
    } else {  // copy the planes of the arg-max element verbatim (hi/lo are a canonical function of the value)
      float* f = static_cast<float*>(cat4.base);
      f[dst] = f[arg[k]];
      if (cat4.format == YP_FMT_F32X2) f[dst + cat4.plane_stride] = f[arg[k] + cat4.plane_stride];
    }
  }
}

}  // namespace
}  // namespace yp

extern "C" int yp_nchw_to_s2d(const float* x, int32_t B, int32_t H, int32_t W, const YpView* out, void* stream) {
  return yp::to_s2d(x, false, B, H, W, out, static_cast<cudaStream_t>(stream));
}

extern "C" int yp_frame_to_s2d(const uint8_t* frame, int32_t B, int32_t H, int32_t W, const YpView* out, void* stream) {
  return yp::to_s2d(frame, true, B, H, W, out, static_cast<cudaStream_t>(stream));
}

extern "C" int yp_nhwc_to_nchw(const YpView* in, int32_t C, float* out, void* stream) {
  YP_REQUIRE(in && in->base && out, YP_ERR_ARG, "nhwc_to_nchw: null pointer");
  YP_REQUIRE(C > 0 && C <= in->C, YP_ERR_SHAPE, "nhwc_to_nchw: C=%d exceeds view channels %d", C, in->C);
  const int64_t HW = static_cast<int64_t>(in->H) * in->W;
  dim3 grid(static_cast<unsigned>(yp::ceil_div64(HW, 32)), yp::ceil_div(C, 32), in->B);
  yp::nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(*in, C, out);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_sppf_pool(const YpView* cat4, void* stream) {
  YP_REQUIRE(cat4 && cat4->base, YP_ERR_ARG, "sppf_pool: null view");
  YP_REQUIRE(cat4->C % 4 == 0, YP_ERR_SHAPE, "sppf_pool: concat buffer channels %d not a multiple of 4", cat4->C);
  const int C = cat4->C / 4;
  const int64_t total = static_cast<int64_t>(cat4->B) * cat4->H * cat4->W * C;
  yp::sppf_pool_kernel<<<static_cast<unsigned>(yp::ceil_div64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*cat4, C);
  YP_LAUNCH_OK();
  return YP_OK;
}
