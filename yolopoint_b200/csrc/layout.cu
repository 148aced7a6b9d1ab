// Layout / pooling kernels: input conversion to the stem's space-to-depth NHWC operand, NHWC -> NCHW
// export of network outputs, and the single-pass SPPF pooling.  All are HBM/L2-bound element movers.
#include <algorithm>

#include "common.cuh"
#include "sppf.cuh"

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

// uint8 HWC frame -> space-to-depth operand, staged through shared memory so that both sides are coalesced: a CTA owns a strip of
// S2D_PIX output pixels of one output row, i.e. two input row segments of 6 * S2D_PIX contiguous bytes, which it reads as aligned
// 32-bit words (W % 4 == 0, 4-byte aligned frame); then consecutive threads emit consecutive 16-byte vectors of the output rows
// (4 fp32 channels per plane, or 8 bf16 channels).  Output channel ch = (ph*2 + pw)*3 + c of pixel px is byte 6*px + ch % 6 of
// input row ph = ch / 6.  Same values as to_s2d_kernel<true>: byte / 255 in round-to-nearest division, taken from a 256-entry table
// the CTA builds once (one IEEE division per thread instead of 12-16 per pixel: ncu showed the kernel bound by instruction issue,
// 88 % issue-slot utilisation, with the divisions' FFMA / FCHK / branch sequences on top).  A thread per output pixel with byte
// loads and 32-/64-byte stores per thread reached 39 % (bf16) / 64 % (fp32 pairs) of the HBM peak.
constexpr int S2D_PIX = 512;

template <int kFmt>
__global__ void __launch_bounds__(256) frame_to_s2d_staged_kernel(const uint8_t* __restrict__ frame, int B, int H, int W, YpView out) {
  __shared__ uint32_t raw[2][S2D_PIX * 6 / 4];
  __shared__ float lut[256];
  lut[threadIdx.x] = __fdiv_rn(static_cast<float>(threadIdx.x), 255.0f);      // blockDim.x == 256
  const int W2 = W / 2, H2 = H / 2;
  const int strips = (W2 + S2D_PIX - 1) / S2D_PIX;
  const int strip = blockIdx.x % strips, row = blockIdx.x / strips;
  const int h2 = row % H2, b = row / H2;
  const int w0 = strip * S2D_PIX;
  const int npx = min(S2D_PIX, W2 - w0);
  const int nwords = npx * 6 / 4;                       // W2 and w0 are even, so npx is
  for (int t = threadIdx.x; t < 2 * nwords; t += blockDim.x) {
    const int ph = t >= nwords ? 1 : 0, k = t - ph * nwords;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(frame + ((static_cast<int64_t>(b) * H + 2 * h2 + ph) * W + 2 * w0) * 3);
    raw[ph][k] = __ldg(src + k);
  }
  __syncthreads();
  const uint8_t* rb = reinterpret_cast<const uint8_t*>(&raw[0][0]);
  const int64_t pix0 = (static_cast<int64_t>(b) * H2 + h2) * W2 + w0;
  auto value = [&](int px, int ch) -> float {          // ch < 12
    const int ph = ch >= 6 ? 1 : 0;
    return lut[rb[ph * (S2D_PIX * 6) + 6 * px + (ch - 6 * ph)]];
  };
  if (kFmt == YP_FMT_BF16) {
    for (int t = threadIdx.x; t < npx * 2; t += blockDim.x) {
      const int px = t >> 1, half = t & 1;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { const int ch = half * 8 + e; v[e] = ch < 12 ? value(px, ch) : 0.0f; }
      uint4 pk;
      __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int e = 0; e < 4; ++e) p2[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out.base) + (pix0 + px) * out.pix_stride + half * 8) = pk;
    }
  } else {
    for (int t = threadIdx.x; t < npx * 4; t += blockDim.x) {
      const int px = t >> 2, q = t & 3;
      float hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ch = q * 4 + e;
        const float x = ch < 12 ? value(px, ch) : 0.0f;
        hi[e] = kFmt == YP_FMT_F32X2 ? tf32_round(x) : x;
        lo[e] = tf32_round(x - hi[e]);
      }
      float* o = static_cast<float*>(out.base) + (pix0 + px) * out.pix_stride + q * 4;
      *reinterpret_cast<float4*>(o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      if (kFmt == YP_FMT_F32X2) *reinterpret_cast<float4*>(o + out.plane_stride) = make_float4(lo[0], lo[1], lo[2], lo[3]);
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
  if (frame && W % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3u) == 0) {
    const int64_t ctas = static_cast<int64_t>(B) * (H / 2) * ceil_div(W / 2, S2D_PIX);
    YP_REQUIRE(ctas < (1ll << 31), YP_ERR_SHAPE, "to_s2d: frame batch too large");
    const uint8_t* f = static_cast<const uint8_t*>(src);
    if (out->format == YP_FMT_BF16) frame_to_s2d_staged_kernel<YP_FMT_BF16><<<static_cast<unsigned>(ctas), 256, 0, st>>>(f, B, H, W, *out);
    else if (out->format == YP_FMT_F32X2) frame_to_s2d_staged_kernel<YP_FMT_F32X2><<<static_cast<unsigned>(ctas), 256, 0, st>>>(f, B, H, W, *out);
    else frame_to_s2d_staged_kernel<YP_FMT_F32><<<static_cast<unsigned>(ctas), 256, 0, st>>>(f, B, H, W, *out);
  } else if (frame) to_s2d_kernel<true><<<blocks, 256, 0, st>>>(src, B, H, W, *out);
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

// Fast path of the export for plain fp32 views with 16-byte aligned pixels: a CTA moves 32 pixels x 64 channels; every lane reads a
// float4 (16 lanes = one pixel's 64 channels, a warp = two pixels), the tile is transposed in shared memory and written back as
// 128-byte pixel runs per channel.  The element-wise kernel above spent ~70 instructions per warp-element on index arithmetic and
// reached 50 % of the HBM peak (instruction issue 80 % busy).
__global__ void __launch_bounds__(256) nhwc_to_nchw_f32_kernel(YpView in, int C, float* __restrict__ out) {
  __shared__ float tile[64][33];
  const int64_t HW = static_cast<int64_t>(in.H) * in.W;
  const int b = blockIdx.z, c0 = blockIdx.y * 64;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* src = static_cast<const float*>(in.base) + (static_cast<int64_t>(b) * HW + p0) * in.pix_stride + c0;
  const int cl = (lane & 15) * 4, ph = lane >> 4;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int p = warp * 4 + it * 2 + ph;                      // pixel of the tile this lane loads
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p0 + p < HW && c0 + cl < in.C) v = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(p) * in.pix_stride + cl));
    tile[cl][p] = v.x; tile[cl + 1][p] = v.y; tile[cl + 2][p] = v.z; tile[cl + 3][p] = v.w;
  }
  __syncthreads();
  float* dst = out + (static_cast<int64_t>(b) * C + c0) * HW + p0 + lane;
  const bool pix_ok = p0 + lane < HW;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = warp * 8 + j;
    if (pix_ok && c0 + c < C) dst[static_cast<int64_t>(c) * HW] = tile[c][lane];
  }
}

// desc / ||desc||_2 over the channels of every pixel of a plain fp32 NHWC view, in place (models/YOLOPoint.py:219-220: the
// normalisation of the descriptor head).  The conv kernel does this in its epilogue when all D channels of a pixel sit in one
// accumulator tile (D <= 256); version "x" (D = 320) runs the last descriptor convolution without it and this kernel after it.
// One warp per pixel, float4 lanes; HBM-bound: reads and writes D floats per pixel.
__global__ void __launch_bounds__(256) l2norm_rows_kernel(YpView v, int64_t n_pix) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nv = v.C >> 2;
  for (int64_t p = warp0; p < n_pix; p += n_warps) {
    float4* row = reinterpret_cast<float4*>(static_cast<float*>(v.base) + p * v.pix_stride);
    float ss = 0.0f;
    for (int j = lane; j < nv; j += 32) {
      const float4 f = row[j];
      ss = fmaf(f.x, f.x, ss); ss = fmaf(f.y, f.y, ss); ss = fmaf(f.z, f.z, ss); ss = fmaf(f.w, f.w, ss);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    const float inv = 1.0f / sqrtf(ss);
    for (int j = lane; j < nv; j += 32) {
      float4 f = row[j];
      f.x *= inv; f.y *= inv; f.z *= inv; f.w *= inv;
      row[j] = f;
    }
  }
}

// fp32 -> (hi, lo) TF32 operand planes (hi = tf32(x), lo = tf32(x - hi); cvt.rna): the operand split of the 3xTF32 tensor-core match
// (descriptor rows become the "pixels" / "weights" of a 1x1 convolution).  One float4 per thread, HBM-bound: 4 B in, 8 B out per element.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float4* __restrict__ src, int64_t n4, float4* __restrict__ hi, float4* __restrict__ lo) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(src + i);
    float4 h, l;
    h.x = tf32_round(v.x); h.y = tf32_round(v.y); h.z = tf32_round(v.z); h.w = tf32_round(v.w);
    l.x = tf32_round(v.x - h.x); l.y = tf32_round(v.y - h.y); l.z = tf32_round(v.z - h.z); l.w = tf32_round(v.w - h.w);
    hi[i] = h; lo[i] = l;
  }
}

// SPPF pooling: see sppf.cuh (the body is shared with conv_chain_kernel, which runs it as an in-chain operation).
__global__ void __launch_bounds__(256) sppf_pool_kernel(YpView cat4, int C) {
  extern __shared__ unsigned char sp_smem[];
  sppf_pool_item(cat4, C, blockIdx.x, blockIdx.y, sp_smem);
}

// MaxPool2d(kernel 2, stride 2) between two NHWC views of one format (YOLOPointv52 descriptor head, src/models/YOLOPoint.py:287, 311).
// One thread owns 8 consecutive channels of one output pixel: 16-byte loads of the four window pixels (per operand plane), the
// winner of each channel is chosen on its value (hi + lo for YP_FMT_F32X2) and its operand planes are copied unchanged, so the
// stored elements are bit-identical to the source's.  Ties keep the first pixel in window raster order.
template <int kFmt>
__global__ void __launch_bounds__(256) maxpool2x2_kernel(YpView in, YpView out, int groups) {
  const int Ho = out.H, Wo = out.W;
  const int64_t total = static_cast<int64_t>(out.B) * Ho * Wo * groups;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    const int64_t op = idx / groups;
    const int wo = static_cast<int>(op % Wo);
    const int ho = static_cast<int>((op / Wo) % Ho);
    const int b = static_cast<int>(op / (static_cast<int64_t>(Wo) * Ho));
    const int64_t dst = op * out.pix_stride + g * 8;
    int64_t srcp[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      srcp[q] = ((static_cast<int64_t>(b) * in.H + 2 * ho + (q >> 1)) * in.W + 2 * wo + (q & 1)) * in.pix_stride + g * 8;
    if (kFmt == YP_FMT_BF16) {
      const __nv_bfloat16* f = static_cast<const __nv_bfloat16*>(in.base);
      uint4 best = *reinterpret_cast<const uint4*>(f + srcp[0]);
      __nv_bfloat16* bb = reinterpret_cast<__nv_bfloat16*>(&best);
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        const uint4 v = *reinterpret_cast<const uint4*>(f + srcp[q]);
        const __nv_bfloat16* vv = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (__bfloat162float(vv[i]) > __bfloat162float(bb[i])) bb[i] = vv[i];
      }
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out.base) + dst) = best;
    } else {
      const float* f = static_cast<const float*>(in.base);
      const bool two = kFmt == YP_FMT_F32X2;
      float hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float h[8], l[8];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(f + srcp[q] + 4 * i);
          h[4 * i] = a.x; h[4 * i + 1] = a.y; h[4 * i + 2] = a.z; h[4 * i + 3] = a.w;
          float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
          if (two) c = *reinterpret_cast<const float4*>(f + in.plane_stride + srcp[q] + 4 * i);
          l[4 * i] = c.x; l[4 * i + 1] = c.y; l[4 * i + 2] = c.z; l[4 * i + 3] = c.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (q == 0 || __fadd_rn(h[i], l[i]) > __fadd_rn(hi[i], lo[i])) { hi[i] = h[i]; lo[i] = l[i]; }
      }
      float* o = static_cast<float*>(out.base);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        *reinterpret_cast<float4*>(o + dst + 4 * i) = make_float4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
        if (two) *reinterpret_cast<float4*>(o + out.plane_stride + dst + 4 * i) = make_float4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
      }
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
  if (in->format == YP_FMT_F32 && in->C % 4 == 0 && in->pix_stride % 4 == 0 && yp::aligned16(in->base)) {
    dim3 grid(static_cast<unsigned>(yp::ceil_div64(HW, 32)), yp::ceil_div(C, 64), in->B);
    yp::nhwc_to_nchw_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*in, C, out);
  } else {
    dim3 grid(static_cast<unsigned>(yp::ceil_div64(HW, 32)), yp::ceil_div(C, 32), in->B);
    yp::nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(*in, C, out);
  }
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_l2norm_nhwc(const YpView* v, void* stream) {
  YP_REQUIRE(v && v->base, YP_ERR_ARG, "l2norm: null view");
  YP_REQUIRE(v->format == YP_FMT_F32 && v->C % 4 == 0 && v->pix_stride % 4 == 0 && yp::aligned16(v->base), YP_ERR_SHAPE,
             "l2norm: needs a plain fp32 view with 16-byte aligned rows (format %d, C %d, pixel stride %lld)", v->format, v->C, (long long)v->pix_stride);
  const int64_t n_pix = static_cast<int64_t>(v->B) * v->H * v->W;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(yp::ceil_div64(n_pix, 8), static_cast<int64_t>(yp::sm_count()) * 8));
  yp::l2norm_rows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*v, n_pix);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_split_tf32(const float* src, int64_t n, float* hi, float* lo, void* stream) {
  YP_REQUIRE(src && hi && lo, YP_ERR_ARG, "split_tf32: null pointer");
  YP_REQUIRE(n >= 0 && n % 4 == 0 && yp::aligned16(src) && yp::aligned16(hi) && yp::aligned16(lo), YP_ERR_ALIGN, "split_tf32: n must be a multiple of 4 and the buffers 16-byte aligned");
  if (n == 0) return YP_OK;
  const int64_t n4 = n / 4;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(yp::ceil_div64(n4, 256), static_cast<int64_t>(yp::sm_count()) * 16));
  yp::split_tf32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(src), n4, reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo));
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_sppf_pool(const YpView* cat4, void* stream) {
  YP_REQUIRE(cat4 && cat4->base, YP_ERR_ARG, "sppf_pool: null view");
  YP_REQUIRE(cat4->C % 4 == 0 && (cat4->C / 4) % yp::SPPF_G == 0, YP_ERR_SHAPE, "sppf_pool: concat buffer channels %d not a multiple of %d", cat4->C, 4 * yp::SPPF_G);
  const int C = cat4->C / 4;
  const int HW = cat4->H * cat4->W;
  YP_REQUIRE(HW <= 65535, YP_ERR_SHAPE, "sppf_pool: feature map %dx%d too large", cat4->H, cat4->W);
  const size_t smem = yp::sppf_smem_bytes(HW);
  YP_REQUIRE(smem <= 200 * 1024, YP_ERR_SHAPE, "sppf_pool: feature map %dx%d needs %zu bytes of shared memory", cat4->H, cat4->W, smem);
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    YP_CUDA_OK(cudaFuncSetAttribute(yp::sppf_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  yp::sppf_pool_kernel<<<dim3(C / yp::SPPF_G, cat4->B), 256, smem, static_cast<cudaStream_t>(stream)>>>(*cat4, C);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_maxpool2x2(const YpView* in, const YpView* out, void* stream) {
  YP_REQUIRE(in && out && in->base && out->base, YP_ERR_ARG, "maxpool2x2: null view");
  YP_REQUIRE(in->format == out->format && (in->format == YP_FMT_BF16 || in->format == YP_FMT_F32X2 || in->format == YP_FMT_F32), YP_ERR_SHAPE,
             "maxpool2x2: formats %d -> %d unsupported", in->format, out->format);
  YP_REQUIRE(in->H % 2 == 0 && in->W % 2 == 0 && out->B == in->B && out->H == in->H / 2 && out->W == in->W / 2 && out->C == in->C && in->C % 8 == 0 &&
                 out->upsample <= 1, YP_ERR_SHAPE, "maxpool2x2: [%d,%d,%d,%d] -> [%d,%d,%d,%d] is not a 2x2 stride-2 pooling of 8-channel groups",
             in->B, in->H, in->W, in->C, out->B, out->H, out->W, out->C);
  const int es = yp::fmt_esize(in->format);
  YP_REQUIRE(yp::aligned16(in->base) && yp::aligned16(out->base) && (in->pix_stride * es) % 16 == 0 && (out->pix_stride * es) % 16 == 0 &&
                 (in->plane_stride * es) % 16 == 0 && (out->plane_stride * es) % 16 == 0, YP_ERR_ALIGN, "maxpool2x2: views not 16-byte aligned");
  const int groups = in->C / 8;
  const int64_t total = static_cast<int64_t>(out->B) * out->H * out->W * groups;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(yp::ceil_div64(total, 256), static_cast<int64_t>(yp::sm_count()) * 16));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in->format == YP_FMT_BF16) yp::maxpool2x2_kernel<YP_FMT_BF16><<<blocks, 256, 0, st>>>(*in, *out, groups);
  else if (in->format == YP_FMT_F32X2) yp::maxpool2x2_kernel<YP_FMT_F32X2><<<blocks, 256, 0, st>>>(*in, *out, groups);
  else yp::maxpool2x2_kernel<YP_FMT_F32><<<blocks, 256, 0, st>>>(*in, *out, groups);
  YP_LAUNCH_OK();
  return YP_OK;
}
