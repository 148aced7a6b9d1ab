// Training-step glue between the convolutions on bf16 NHWC activations (SURVEY.md section 8 row a11 / 8b `sppf_fwd/bwd`):
//   * channel concatenation with the resampling the network applies to a source on the way in, forward and backward in one launch each:
//       copy               torch.cat                                  (C3 / heads, src/models/common.py:123-135)
//       nearest 2x         nn.Upsample(scale_factor=2) -> cat         (src/models/YOLOPoint.py:214, 222-223)
//       2x2 max pool       nn.MaxPool2d(2, 2) -> cat                  (YOLOPointv52, src/models/YOLOPoint.py:311)
//     backward = split of the gradient with the adjoint resampling (sum of the 4 children / routing to the first maximum of the window).
//   * SPPF (src/models/common.py:213-229): x -> cat(x, m(x), m(m(x)), m(m(m(x)))) with m = MaxPool2d(5, 1, 2), forward with the source
//     pixel of every pooled value recorded, backward = one scatter of the three pooled gradients to those pixels (fp32 sums in shared
//     memory), which is what the three chained max_pool2d backward passes of autograd compute.
// PyTorch runs these as 1 + n launches forward (upsample / pool, cat) and one strided copy per consumer backward; here a thread owns
// one 16-byte vector (8 channels) of one pixel, every access is a coalesced 16-byte load / store.  HBM-bound streaming kernels.
#include <algorithm>

#include "common.cuh"

namespace yp {
namespace {

constexpr int kCatMaxParts = YP_CAT_MAX_PARTS;

struct CatArgs {
  const uint4* src[kCatMaxParts];
  uint4* grad[kCatMaxParts];
  int groups[kCatMaxParts];      // C / 8
  int mode[kCatMaxParts];
  int goff[kCatMaxParts + 1];    // prefix sums of groups (forward: position in the output pixel)
  int64_t voff[kCatMaxParts + 1];   // prefix sums of source vectors (backward: flat index over all sources)
  int n, B, H, W;
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 t = __bfloat1622float2(h[e]); f[2 * e] = t.x; f[2 * e + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return u;
}

__device__ __forceinline__ void src_dims(const CatArgs& a, int part, int* Hs, int* Ws) {
  *Hs = a.mode[part] == YP_CAT_UP2 ? a.H / 2 : a.mode[part] == YP_CAT_POOL2 ? a.H * 2 : a.H;
  *Ws = a.mode[part] == YP_CAT_UP2 ? a.W / 2 : a.mode[part] == YP_CAT_POOL2 ? a.W * 2 : a.W;
}

constexpr int kCatUnroll = 4;   // independent 16-byte vectors a thread keeps in flight

// value of output vector `idx` (pixel-major, G vectors per pixel).  Idx = uint32_t whenever the tensors allow it: the index arithmetic
// (two divisions per vector) is then 32-bit -- 64-bit divisions made the first streaming kernels of this library issue-bound.
template <typename Idx>
__device__ __forceinline__ uint4 cat_fwd_value(const CatArgs& a, int G, Idx idx) {
  const int g = static_cast<int>(idx % static_cast<Idx>(G));
  const Idx pix = idx / static_cast<Idx>(G);
  int part = 0;
  while (part + 1 < a.n && g >= a.goff[part + 1]) ++part;
  const int gl = g - a.goff[part], Gs = a.groups[part];
  const uint4* s = a.src[part];
  if (a.mode[part] == YP_CAT_COPY) return __ldg(s + static_cast<int64_t>(pix) * Gs + gl);
  const int w = static_cast<int>(pix % static_cast<Idx>(a.W));
  const Idx row = pix / static_cast<Idx>(a.W);
  const int h = static_cast<int>(row % static_cast<Idx>(a.H)), b = static_cast<int>(row / static_cast<Idx>(a.H));
  if (a.mode[part] == YP_CAT_UP2) return __ldg(s + ((static_cast<int64_t>(b) * (a.H / 2) + h / 2) * (a.W / 2) + w / 2) * Gs + gl);
  // first maximum of the 2x2 window in raster order (max_pool2d)
  const int Ws = 2 * a.W;
  const int64_t p00 = (static_cast<int64_t>(b) * 2 * a.H + 2 * h) * Ws + 2 * w;
  uint4 u[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) u[q] = __ldg(s + (p00 + (q >> 1) * Ws + (q & 1)) * Gs + gl);
  __nv_bfloat16* bb = reinterpret_cast<__nv_bfloat16*>(&u[0]);
#pragma unroll
  for (int q = 1; q < 4; ++q) {
    const __nv_bfloat16* uu = reinterpret_cast<const __nv_bfloat16*>(&u[q]);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (__bfloat162float(uu[e]) > __bfloat162float(bb[e])) bb[e] = uu[e];
  }
  return u[0];
}

template <typename Idx>
__global__ void __launch_bounds__(256) cat_fwd_kernel(const CatArgs a, uint4* __restrict__ out, Idx total) {
  const int G = a.goff[a.n];
  const Idx stride = static_cast<Idx>(gridDim.x) * blockDim.x;
  for (Idx idx = blockIdx.x * static_cast<Idx>(blockDim.x) + threadIdx.x; idx < total; idx += kCatUnroll * stride) {
    uint4 v[kCatUnroll];
#pragma unroll
    for (int k = 0; k < kCatUnroll; ++k)
      if (idx + k * stride < total) v[k] = cat_fwd_value<Idx>(a, G, idx + k * stride);
#pragma unroll
    for (int k = 0; k < kCatUnroll; ++k)
      if (idx + k * stride < total) out[idx + k * stride] = v[k];
  }
}

// gradient vector `idx` of the flat list of all source vectors -> its part / position; returns false where the part wants no gradient
template <typename Idx>
__device__ __forceinline__ bool cat_bwd_value(const CatArgs& a, const uint4* __restrict__ dout, int G, Idx idx, int* part_out, Idx* local_out, uint4* r) {
  int part = 0;
  while (part + 1 < a.n && static_cast<int64_t>(idx) >= a.voff[part + 1]) ++part;
  if (a.grad[part] == nullptr) return false;
  const Idx local = idx - static_cast<Idx>(a.voff[part]);
  const int Gs = a.groups[part], gl = static_cast<int>(local % static_cast<Idx>(Gs)), g = a.goff[part] + gl;
  const Idx spix = local / static_cast<Idx>(Gs);
  *part_out = part;
  *local_out = local;
  if (a.mode[part] == YP_CAT_COPY) {
    *r = __ldg(dout + static_cast<int64_t>(spix) * G + g);
    return true;
  }
  int Hs, Ws;
  src_dims(a, part, &Hs, &Ws);
  const int ws = static_cast<int>(spix % static_cast<Idx>(Ws));
  const Idx row = spix / static_cast<Idx>(Ws);
  const int hs = static_cast<int>(row % static_cast<Idx>(Hs)), b = static_cast<int>(row / static_cast<Idx>(Hs));
  if (a.mode[part] == YP_CAT_UP2) {   // sum of the four children in fp32 (upsample_nearest2d backward)
    const int64_t p00 = (static_cast<int64_t>(b) * a.H + 2 * hs) * a.W + 2 * ws;
    uint4 u[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) u[q] = __ldg(dout + (p00 + (q >> 1) * a.W + (q & 1)) * G + g);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, f[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      unpack8(u[q], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
    *r = pack8(acc);
    return true;
  }
  // 2x2 pooling: this source pixel receives the window's gradient where it is the first maximum of its window
  const int q_me = (hs & 1) * 2 + (ws & 1);
  const int64_t p00 = (static_cast<int64_t>(b) * Hs + (hs & ~1)) * Ws + (ws & ~1);
  uint4 u[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) u[q] = __ldg(a.src[part] + (p00 + (q >> 1) * Ws + (q & 1)) * Gs + gl);
  const uint4 du = __ldg(dout + ((static_cast<int64_t>(b) * a.H + hs / 2) * a.W + ws / 2) * G + g);
  float win[4][8], d[8], o[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) unpack8(u[q], win[q]);
  unpack8(du, d);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    int arg = 0;
    float best = win[0][e];
#pragma unroll
    for (int q = 1; q < 4; ++q)
      if (win[q][e] > best) { best = win[q][e]; arg = q; }
    o[e] = arg == q_me ? d[e] : 0.f;
  }
  *r = pack8(o);
  return true;
}

template <typename Idx>
__global__ void __launch_bounds__(256) cat_bwd_kernel(const CatArgs a, const uint4* __restrict__ dout, Idx total) {
  const int G = a.goff[a.n];
  const Idx stride = static_cast<Idx>(gridDim.x) * blockDim.x;
  for (Idx idx = blockIdx.x * static_cast<Idx>(blockDim.x) + threadIdx.x; idx < total; idx += kCatUnroll * stride) {
    uint4 v[kCatUnroll];
    int part[kCatUnroll];
    Idx local[kCatUnroll];
    bool ok[kCatUnroll];
#pragma unroll
    for (int k = 0; k < kCatUnroll; ++k) ok[k] = idx + k * stride < total && cat_bwd_value<Idx>(a, dout, G, idx + k * stride, &part[k], &local[k], &v[k]);
#pragma unroll
    for (int k = 0; k < kCatUnroll; ++k)
      if (ok[k]) a.grad[part[k]][local[k]] = v[k];
  }
}

// ---- SPPF ------------------------------------------------------------------------------------------------------------
constexpr int kSpG = 8;   // channels per work item (one 16-byte vector per pixel)

static inline size_t sppf_train_smem(int HW) { return static_cast<size_t>(HW) * kSpG * (2 * sizeof(float) + 2 * sizeof(unsigned short)); }

// item = (channel group, image): x [B,HW,C] -> out [B,HW,4C] (slices x, y1, y2, y3), arg [3][B][HW][C] = source pixel of y_k
__global__ void __launch_bounds__(256) sppf_train_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, unsigned short* __restrict__ arg,
                                                             int B, int H, int W, int C) {
  extern __shared__ unsigned char sp_smem[];
  const int HW = H * W, n = HW * kSpG;
  float* val[2] = {reinterpret_cast<float*>(sp_smem), reinterpret_cast<float*>(sp_smem) + n};
  unsigned short* src[2] = {reinterpret_cast<unsigned short*>(val[1] + n), reinterpret_cast<unsigned short*>(val[1] + n) + n};
  const int c0 = blockIdx.x * kSpG, b = blockIdx.y;
  const int64_t img = static_cast<int64_t>(b) * HW;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int p = i / kSpG, g = i - p * kSpG;
    const __nv_bfloat16 v = x[(img + p) * C + c0 + g];
    val[0][i] = __bfloat162float(v);
    src[0][i] = static_cast<unsigned short>(p);
    out[(img + p) * 4 * C + c0 + g] = v;
  }
  __syncthreads();
  for (int pass = 0; pass < 3; ++pass) {
    const float* vi = val[pass & 1];
    const unsigned short* si = src[pass & 1];
    float* vo = val[(pass + 1) & 1];
    unsigned short* so = src[(pass + 1) & 1];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int p = i / kSpG, g = i - p * kSpG;
      const int y = p / W, xx0 = p - y * W;
      float best = -INFINITY;
      unsigned short a = static_cast<unsigned short>(p);
      for (int yy = max(0, y - 2); yy <= min(H - 1, y + 2); ++yy)
        for (int xx = max(0, xx0 - 2); xx <= min(W - 1, xx0 + 2); ++xx) {
          const int q = (yy * W + xx) * kSpG + g;
          const float v = vi[q];
          if (v > best) { best = v; a = si[q]; }
        }
      vo[i] = best;
      so[i] = a;
      out[(img + p) * 4 * C + static_cast<int64_t>(pass + 1) * C + c0 + g] = __float2bfloat16_rn(best);   // exact: best is a bf16 value
      arg[((static_cast<int64_t>(pass) * B + b) * HW + p) * C + c0 + g] = a;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) sppf_train_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const unsigned short* __restrict__ arg,
                                                             __nv_bfloat16* __restrict__ dx, int B, int H, int W, int C) {
  extern __shared__ unsigned char sp_smem[];
  const int HW = H * W, n = HW * kSpG;
  float* acc = reinterpret_cast<float*>(sp_smem);
  const int c0 = blockIdx.x * kSpG, b = blockIdx.y;
  const int64_t img = static_cast<int64_t>(b) * HW;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int p = i / kSpG, g = i - p * kSpG;
    acc[i] = __bfloat162float(dout[(img + p) * 4 * C + c0 + g]);
  }
  __syncthreads();
  for (int pass = 0; pass < 3; ++pass)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int p = i / kSpG, g = i - p * kSpG;
      const int a = arg[((static_cast<int64_t>(pass) * B + b) * HW + p) * C + c0 + g];
      atomicAdd(acc + a * kSpG + g, __bfloat162float(dout[(img + p) * 4 * C + static_cast<int64_t>(pass + 1) * C + c0 + g]));
    }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int p = i / kSpG, g = i - p * kSpG;
    dx[(img + p) * C + c0 + g] = __float2bfloat16_rn(acc[i]);
  }
}

int fill_cat_args(const YpCatPart* parts, int n, int B, int H, int W, bool backward, CatArgs* a) {
  YP_REQUIRE(parts && n >= 1 && n <= kCatMaxParts && B > 0 && H > 0 && W > 0, YP_ERR_ARG, "cat: n=%d B=%d H=%d W=%d", n, B, H, W);
  memset(a, 0, sizeof(*a));
  a->n = n; a->B = B; a->H = H; a->W = W;
  for (int i = 0; i < n; ++i) {
    const YpCatPart& p = parts[i];
    YP_REQUIRE(p.C > 0 && p.C % 8 == 0, YP_ERR_SHAPE, "cat: part %d has %d channels (multiples of 8 only)", i, p.C);
    YP_REQUIRE(p.mode == YP_CAT_COPY || p.mode == YP_CAT_UP2 || p.mode == YP_CAT_POOL2, YP_ERR_ARG, "cat: part %d: mode %d", i, p.mode);
    YP_REQUIRE(p.mode != YP_CAT_UP2 || (H % 2 == 0 && W % 2 == 0), YP_ERR_SHAPE, "cat: part %d: 2x upsampling into an odd %dx%d map", i, H, W);
    const bool need_src = !backward || p.mode == YP_CAT_POOL2;
    YP_REQUIRE(!need_src || (p.src && aligned16(p.src)), YP_ERR_ALIGN, "cat: part %d: source missing or not 16-byte aligned", i);
    YP_REQUIRE(!backward || p.grad == nullptr || aligned16(p.grad), YP_ERR_ALIGN, "cat: part %d: gradient not 16-byte aligned", i);
    a->src[i] = static_cast<const uint4*>(p.src);
    a->grad[i] = static_cast<uint4*>(p.grad);
    a->groups[i] = p.C / 8;
    a->mode[i] = p.mode;
    a->goff[i + 1] = a->goff[i] + p.C / 8;
    const int64_t Hs = p.mode == YP_CAT_UP2 ? H / 2 : p.mode == YP_CAT_POOL2 ? 2 * H : H, Ws = p.mode == YP_CAT_UP2 ? W / 2 : p.mode == YP_CAT_POOL2 ? 2 * W : W;
    a->voff[i + 1] = a->voff[i] + static_cast<int64_t>(B) * Hs * Ws * (p.C / 8);
  }
  for (int i = n; i < kCatMaxParts; ++i) { a->goff[i + 1] = a->goff[n]; a->voff[i + 1] = a->voff[n]; }
  return YP_OK;
}

unsigned stream_blocks(int64_t total) { return static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(total, 256), static_cast<int64_t>(sm_count()) * 16))); }

}  // namespace
}  // namespace yp

extern "C" int yp_cat_nhwc_fwd(const YpCatPart* parts, int32_t n, void* out, int32_t B, int32_t H, int32_t W, void* stream) {
  yp::CatArgs a;
  const int rc = yp::fill_cat_args(parts, n, B, H, W, false, &a);
  if (rc != YP_OK) return rc;
  YP_REQUIRE(out && yp::aligned16(out), YP_ERR_ALIGN, "cat_fwd: output missing or not 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * H * W * a.goff[n];
  const unsigned blocks = yp::stream_blocks(yp::ceil_div64(total, yp::kCatUnroll));
  if (total < (int64_t(1) << 31))   // (index + unroll * grid stride stays below 2^32)
    yp::cat_fwd_kernel<uint32_t><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, static_cast<uint4*>(out), static_cast<uint32_t>(total));
  else
    yp::cat_fwd_kernel<int64_t><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, static_cast<uint4*>(out), total);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_cat_nhwc_bwd(const YpCatPart* parts, int32_t n, const void* dout, int32_t B, int32_t H, int32_t W, void* stream) {
  yp::CatArgs a;
  const int rc = yp::fill_cat_args(parts, n, B, H, W, true, &a);
  if (rc != YP_OK) return rc;
  YP_REQUIRE(dout && yp::aligned16(dout), YP_ERR_ALIGN, "cat_bwd: output gradient missing or not 16-byte aligned");
  const int64_t total = a.voff[n];
  const unsigned blocks = yp::stream_blocks(yp::ceil_div64(total, yp::kCatUnroll));
  if (total < (int64_t(1) << 31))
    yp::cat_bwd_kernel<uint32_t><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, static_cast<const uint4*>(dout), static_cast<uint32_t>(total));
  else
    yp::cat_bwd_kernel<int64_t><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, static_cast<const uint4*>(dout), total);
  YP_LAUNCH_OK();
  return YP_OK;
}

static int sppf_train_check(const char* what, int32_t B, int32_t H, int32_t W, int32_t C, size_t* smem, const void* kernel, bool* configured) {
  YP_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % yp::kSpG == 0, YP_ERR_SHAPE, "%s: B=%d H=%d W=%d C=%d (C must be a multiple of %d)", what, B, H, W, C, yp::kSpG);
  YP_REQUIRE(H * W <= 65535, YP_ERR_SHAPE, "%s: feature map %dx%d too large", what, H, W);
  *smem = yp::sppf_train_smem(H * W);
  YP_REQUIRE(*smem <= 200 * 1024, YP_ERR_SHAPE, "%s: feature map %dx%d needs %zu bytes of shared memory", what, H, W, *smem);
  if (*smem > 48 * 1024 && !*configured) {
    YP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    *configured = true;
  }
  return YP_OK;
}

extern "C" int yp_sppf_train_fwd(const void* x, void* out4, uint16_t* arg, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  YP_REQUIRE(x && out4 && arg, YP_ERR_ARG, "sppf_train_fwd: null pointer");
  size_t smem = 0;
  static thread_local bool configured = false;
  const int rc = sppf_train_check("sppf_train_fwd", B, H, W, C, &smem, reinterpret_cast<const void*>(yp::sppf_train_fwd_kernel), &configured);
  if (rc != YP_OK) return rc;
  yp::sppf_train_fwd_kernel<<<dim3(C / yp::kSpG, B), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out4), arg, B, H, W, C);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_sppf_train_bwd(const void* dout4, const uint16_t* arg, void* dx, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  YP_REQUIRE(dout4 && arg && dx, YP_ERR_ARG, "sppf_train_bwd: null pointer");
  size_t smem = 0;
  static thread_local bool configured = false;
  const int rc = sppf_train_check("sppf_train_bwd", B, H, W, C, &smem, reinterpret_cast<const void*>(yp::sppf_train_bwd_kernel), &configured);
  if (rc != YP_OK) return rc;
  yp::sppf_train_bwd_kernel<<<dim3(C / yp::kSpG, B), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout4), arg, static_cast<__nv_bfloat16*>(dx), B, H, W, C);
  YP_LAUNCH_OK();
  return YP_OK;
}
