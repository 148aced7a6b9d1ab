// Descriptor sampling (bilinear grid_sample + L2 normalise) and brute-force two-way matching
// (demo.py:200-215, 300-341; evaluations/descriptor_evaluation.py:148-181 of the reference).
#include <algorithm>

#include "common.cuh"

namespace yp {
namespace {

// ------------------------------------------------------------------------------------------------
// sampling: one warp per keypoint, lanes stride over the D channels
// ------------------------------------------------------------------------------------------------
constexpr int kMaxDPerLane = 16;  // D <= 512

// K = channel blocks of 32 per descriptor (compile time, so that all 4*K corner loads are issued before the first use; K = 0: run
// time, D <= 512).  The first version loaded inside `if (corner in bounds)` regions: the compiler kept every load in its own branch
// region and each warp paid up to 24 serialised DRAM round trips (ncu: all stall samples on the multiply after each load, 31 % of
// the HBM peak).  Out-of-bounds corners (zeros padding) now read a clamped, valid address and get weight 0: adding +-0 does not
// change the sum, so the result is bit-identical to skipping the term.
template <int K>
__global__ void __launch_bounds__(256) sample_desc_kernel(const float* __restrict__ desc, int B, int D, int Hc, int Wc, long long sB, long long sD,
                                                           long long sH, long long sW, int img_h, int img_w, const float* __restrict__ pts,
                                                           const int* __restrict__ count, int pts_ld, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (wid >= static_cast<int64_t>(B) * pts_ld) return;
  const int b = static_cast<int>(wid / pts_ld), i = static_cast<int>(wid - static_cast<int64_t>(b) * pts_ld);
  int n = count ? count[b] : pts_ld;
  if (n > pts_ld) n = pts_ld;
  if (i >= n) return;
  const float* p = pts + (static_cast<int64_t>(b) * pts_ld + i) * 3;
  // demo.py:207-211: normalise in float64, cast to float32; ATen grid_sampler(align_corners=True): ((g+1)/2)*(size-1)
  const float gx = static_cast<float>(static_cast<double>(p[0]) / (static_cast<double>(img_w) / 2.0) - 1.0);
  const float gy = static_cast<float>(static_cast<double>(p[1]) / (static_cast<double>(img_h) / 2.0) - 1.0);
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), static_cast<float>(Wc - 1));
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), static_cast<float>(Hc - 1));
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const float fx1 = __fadd_rn(fx0, 1.0f), fy1 = __fadd_rn(fy0, 1.0f);
  const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0), x1 = x0 + 1, y1 = y0 + 1;
  const bool okx0 = x0 >= 0 && x0 < Wc, okx1 = x1 >= 0 && x1 < Wc, oky0 = y0 >= 0 && y0 < Hc, oky1 = y1 >= 0 && y1 < Hc;
  const float wnw = (okx0 && oky0) ? __fmul_rn(__fsub_rn(fx1, ix), __fsub_rn(fy1, iy)) : 0.0f;
  const float wne = (okx1 && oky0) ? __fmul_rn(__fsub_rn(ix, fx0), __fsub_rn(fy1, iy)) : 0.0f;
  const float wsw = (okx0 && oky1) ? __fmul_rn(__fsub_rn(fx1, ix), __fsub_rn(iy, fy0)) : 0.0f;
  const float wse = (okx1 && oky1) ? __fmul_rn(__fsub_rn(ix, fx0), __fsub_rn(iy, fy0)) : 0.0f;
  const int cx0 = min(max(x0, 0), Wc - 1), cx1 = min(max(x1, 0), Wc - 1), cy0 = min(max(y0, 0), Hc - 1), cy1 = min(max(y1, 0), Hc - 1);
  const float* base = desc + b * sB;
  const float* pnw = base + cy0 * sH + cx0 * sW;
  const float* pne = base + cy0 * sH + cx1 * sW;
  const float* psw = base + cy1 * sH + cx0 * sW;
  const float* pse = base + cy1 * sH + cx1 * sW;
  constexpr int KK = K > 0 ? K : kMaxDPerLane;
  float c[KK][4];
#pragma unroll
  for (int k = 0; k < KK; ++k) {
    const int d = min(k * 32 + lane, D - 1);                       // lanes / blocks past D re-read the last channel, result unused
    if (K > 0 || k * 32 < D) {
      const long long off = d * sD;
      c[k][0] = __ldg(pnw + off); c[k][1] = __ldg(pne + off); c[k][2] = __ldg(psw + off); c[k][3] = __ldg(pse + off);
    }
  }
  float v[KK];
  float ss = 0.0f;
#pragma unroll
  for (int k = 0; k < KK; ++k) {
    v[k] = 0.0f;
    if ((K > 0 || k * 32 < D) && k * 32 + lane < D) {
      float acc = __fadd_rn(0.0f, __fmul_rn(c[k][0], wnw));
      acc = __fadd_rn(acc, __fmul_rn(c[k][1], wne));
      acc = __fadd_rn(acc, __fmul_rn(c[k][2], wsw));
      acc = __fadd_rn(acc, __fmul_rn(c[k][3], wse));
      v[k] = acc;
      ss = fmaf(acc, acc, ss);
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
  const float nrm = sqrtf(ss);
  float* o = out + (static_cast<int64_t>(b) * pts_ld + i) * D;
#pragma unroll
  for (int k = 0; k < KK; ++k) {
    const int d = k * 32 + lane;
    if ((K > 0 || k * 32 < D) && d < D) o[d] = __fdiv_rn(v[k], nrm);
  }
}

// ------------------------------------------------------------------------------------------------
// matching: 64x64 tiles of d1 . d2^T in fp32 FMA, fused distance + packed-key row/column minima.
// key = float_bits(dist) << 32 | index; dist >= 0 so the bit pattern orders like the value and the low word
// gives numpy.argmin's first-index tie-break.  The N1 x N2 matrix is never written.
// ------------------------------------------------------------------------------------------------
constexpr int TM = 64, TN = 64, TK = 16;

__global__ void match_init_kernel(unsigned long long* row_key, int n1_cap, unsigned long long* col_key, int n2_cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n1_cap) row_key[i] = ~0ull;
  if (i < n2_cap) col_key[i] = ~0ull;
}

// One 64x64 tile of d1 . d2^T (rows i0.., columns j0..) with fused distances and minima.  sel1 / sel2 (optional): row r of the
// operand is row sel[r] of the buffer (the whole-frame pipeline matches the box-filtered subset of the sampled descriptors
// without compacting them first).
__device__ __forceinline__ void match_tile(const float* __restrict__ d1, const int* __restrict__ sel1, int n1, const float* __restrict__ d2,
                                           const int* __restrict__ sel2, int n2, int D, int col_off, int i0, int j0,
                                           unsigned long long* __restrict__ row_key, unsigned long long* __restrict__ col_key,
                                           float (*As)[TM + 4], float (*Bs)[TN + 4], unsigned long long* rmin, unsigned long long* cmin) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.0f;
  if (threadIdx.x < TM) rmin[threadIdx.x] = ~0ull;
  if (threadIdx.x < TN) cmin[threadIdx.x] = ~0ull;
  const int r = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
  const bool va_ok = i0 + r < n1, vb_ok = j0 + r < n2;
  const float* pa = d1 + static_cast<int64_t>(va_ok ? (sel1 ? sel1[i0 + r] : i0 + r) : 0) * D + kq;
  const float* pb = d2 + static_cast<int64_t>(vb_ok ? (sel2 ? sel2[j0 + r] : j0 + r) : 0) * D + kq;
  // 64 rows x 16 k per step: each thread loads one float4 of A and one of B (rows are D-contiguous).  The loads of step k+1 are
  // issued before the FMAs of step k (the first version waited for an L2 round trip in front of every step).
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 va = (va_ok && kq < D) ? *reinterpret_cast<const float4*>(pa) : zero4;
  float4 vb = (vb_ok && kq < D) ? *reinterpret_cast<const float4*>(pb) : zero4;
  for (int k0 = 0; k0 < D; k0 += TK) {
    As[kq][r] = va.x; As[kq + 1][r] = va.y; As[kq + 2][r] = va.z; As[kq + 3][r] = va.w;
    Bs[kq][r] = vb.x; Bs[kq + 1][r] = vb.y; Bs[kq + 2][r] = vb.z; Bs[kq + 3][r] = vb.w;
    __syncthreads();
    const int kn = k0 + TK + kq;
    va = (va_ok && kn < D) ? *reinterpret_cast<const float4*>(pa + k0 + TK) : zero4;
    vb = (vb_ok && kn < D) ? *reinterpret_cast<const float4*>(pb + k0 + TK) : zero4;
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
    }
    __syncthreads();
  }
  // distances + tile-local minima
  unsigned long long rbest[4] = {~0ull, ~0ull, ~0ull, ~0ull}, cbest[4] = {~0ull, ~0ull, ~0ull, ~0ull};
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + ty * 4 + a, j = j0 + tx * 4 + c;
      if (i < n1 && j < n2) {
        const float d = fminf(fmaxf(acc[a][c], -1.0f), 1.0f);
        const float dist = sqrtf(__fsub_rn(2.0f, __fmul_rn(2.0f, d)));
        const unsigned long long hi = static_cast<unsigned long long>(__float_as_uint(dist)) << 32;
        rbest[a] = min(rbest[a], hi | static_cast<unsigned int>(j + col_off));
        cbest[c] = min(cbest[c], hi | static_cast<unsigned int>(i));
      }
    }
#pragma unroll
  for (int a = 0; a < 4; ++a) atomicMin(&rmin[ty * 4 + a], rbest[a]);
#pragma unroll
  for (int c = 0; c < 4; ++c) atomicMin(&cmin[tx * 4 + c], cbest[c]);
  __syncthreads();
  if (threadIdx.x < TM && i0 + threadIdx.x < n1) atomicMin(&row_key[i0 + threadIdx.x], rmin[threadIdx.x]);
  if (threadIdx.x >= TM && threadIdx.x < TM + TN && j0 + (threadIdx.x - TM) < n2) atomicMin(&col_key[j0 + threadIdx.x - TM], cmin[threadIdx.x - TM]);
  __syncthreads();
}

__global__ void __launch_bounds__(256) match_tile_kernel(const float* __restrict__ d1, const int* __restrict__ n1p, int n1_cap,
                                                          const float* __restrict__ d2, const int* __restrict__ n2p, int n2_cap,
                                                          int D, int col_off, unsigned long long* __restrict__ row_key,
                                                          unsigned long long* __restrict__ col_key) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  __shared__ unsigned long long rmin[TM];
  __shared__ unsigned long long cmin[TN];
  const int n1 = n1p ? min(*n1p, n1_cap) : n1_cap;
  const int n2 = n2p ? min(*n2p, n2_cap) : n2_cap;
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  if (i0 >= n1 || j0 >= n2) return;
  match_tile(d1, nullptr, n1, d2, nullptr, n2, D, col_off, i0, j0, row_key, col_key, As, Bs, rmin, cmin);
}

// Whole-frame pipeline form: blockIdx.y = image, a fixed number of CTAs per image walk the tiles the device-side counts call for
// (the grid does not depend on the buffer capacity), operands are rows sel[r] of the per-image descriptor buffers, keys were
// initialised by kp_filter_kernel.
__global__ void __launch_bounds__(256) match_frames_kernel(const float* __restrict__ d1, const int* __restrict__ sel1, const int* __restrict__ n1p,
                                                            const float* __restrict__ d2, const int* __restrict__ sel2, const int* __restrict__ n2p,
                                                            int cap, int D, unsigned long long* __restrict__ row_key,
                                                            unsigned long long* __restrict__ col_key) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  __shared__ unsigned long long rmin[TM];
  __shared__ unsigned long long cmin[TN];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int b = blockIdx.y;
  const int n1 = min(max(n1p[b], 0), cap), n2 = min(max(n2p[b], 0), cap);
  const int nt2 = (n2 + TN - 1) / TN, tiles = ((n1 + TM - 1) / TM) * nt2;
  const int64_t off = static_cast<int64_t>(b) * cap;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int ti = t / nt2, tj = t - ti * nt2;
    match_tile(d1 + off * D, sel1 ? sel1 + off : nullptr, n1, d2 + off * D, sel2 ? sel2 + off : nullptr, n2, D, 0, ti * TM, tj * TN,
               row_key + off, col_key + off, As, Bs, rmin, cmin);
  }
}

// rows sel[i] (i < count) of src -> consecutive rows of dst; one warp per row, float4 lanes (D % 4 == 0)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ sel, const int* __restrict__ count,
                                                           int cap, int D, float* __restrict__ dst) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int n = min(max(count[b], 0), cap);
  const int64_t off = static_cast<int64_t>(b) * cap;
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += gridDim.x * (blockDim.x >> 5)) {
    const float4* s = reinterpret_cast<const float4*>(src + (off + sel[off + i]) * D);
    float4* d = reinterpret_cast<float4*>(dst + (off + i) * D);
    for (int k = lane; k < D / 4; k += 32) d[k] = __ldg(s + k);
  }
}

// ascending-i compaction of mutual nearest neighbours below the threshold (single block)
// blockIdx.x = image (strides img_keys / img_matches elements; 0 for the single-image API).  counts3 (optional, [3][B]): the
// frame's result counts (keypoints, boxes, matches) gathered into one array for a single read-back.
__global__ void match_finalize_kernel(const unsigned long long* __restrict__ row_key, const int* __restrict__ n1p, int n1_cap,
                                      const unsigned long long* __restrict__ col_key, int n2_total, float thr,
                                      float* __restrict__ matches, int* __restrict__ match_count, long long img_keys, long long img_matches,
                                      const int* __restrict__ kcount, const int* __restrict__ bcount, int* __restrict__ counts3) {
  __shared__ int warp_excl[32];
  __shared__ int block_total;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int bimg = blockIdx.x, nimg = gridDim.x;
  row_key += bimg * img_keys; col_key += bimg * img_keys; matches += bimg * img_matches; match_count += bimg;
  if (n1p) n1p += bimg;
  const int n1 = n1p ? min(max(*n1p, 0), n1_cap) : n1_cap;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int carry = 0;
  for (int base = 0; base < n1; base += blockDim.x) {
    const int i = base + threadIdx.x;
    bool keep = false;
    unsigned int j = 0;
    float dist = 0.0f;
    if (i < n1) {
      const unsigned long long k = row_key[i];
      j = static_cast<unsigned int>(k & 0xffffffffull);
      dist = __uint_as_float(static_cast<unsigned int>(k >> 32));
      if (k != ~0ull && j < static_cast<unsigned int>(n2_total))
        keep = dist < thr && static_cast<unsigned int>(col_key[j] & 0xffffffffull) == static_cast<unsigned int>(i);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int incl_w = __popc(m);
    __syncthreads();
    if (lane == 0) warp_excl[wid] = incl_w;
    __syncthreads();
    if (wid == 0) {
      const int wv = lane < (blockDim.x >> 5) ? warp_excl[lane] : 0;
      int wincl = wv;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wincl, s); if (lane >= s) wincl += t; }
      warp_excl[lane] = wincl - wv;
      if (lane == 31) block_total = wincl;
    }
    __syncthreads();
    if (keep) {
      const int pos = carry + warp_excl[wid] + __popc(m & ((1u << lane) - 1u));
      matches[pos * 3] = static_cast<float>(i);
      matches[pos * 3 + 1] = static_cast<float>(j);
      matches[pos * 3 + 2] = dist;
    }
    carry += block_total;
  }
  if (threadIdx.x == 0) {
    *match_count = carry;
    if (counts3) { counts3[bimg] = kcount[bimg]; counts3[nimg + bimg] = bcount[bimg]; counts3[2 * nimg + bimg] = carry; }
  }
}

}  // namespace
}  // namespace yp

extern "C" int yp_sample_desc(const float* desc, int32_t B, int32_t D, int32_t Hc, int32_t Wc, int64_t sB, int64_t sD, int64_t sH,
                              int64_t sW, int32_t img_h, int32_t img_w, const float* pts, const int32_t* count, int32_t pts_ld,
                              float* out, void* stream) {
  YP_REQUIRE(desc && pts && out, YP_ERR_ARG, "sample_desc: null pointer");
  YP_REQUIRE(B > 0 && D > 0 && D <= 32 * yp::kMaxDPerLane && Hc > 0 && Wc > 0 && pts_ld > 0, YP_ERR_SHAPE, "sample_desc: bad shape (D=%d)", D);
  const int64_t warps = static_cast<int64_t>(B) * pts_ld;
  const unsigned blocks = static_cast<unsigned>(yp::ceil_div64(warps * 32, 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define YP_SAMPLE(KB) yp::sample_desc_kernel<KB><<<blocks, 256, 0, st>>>(desc, B, D, Hc, Wc, sB, sD, sH, sW, img_h, img_w, pts, count, pts_ld, out)
  switch ((D + 31) / 32) {   // descriptor widths of the model family: 64 / 128 / 192 / 256 (N / S / M / L), 320 (X)
    case 1: YP_SAMPLE(1); break;
    case 2: YP_SAMPLE(2); break;
    case 4: YP_SAMPLE(4); break;
    case 6: YP_SAMPLE(6); break;
    case 8: YP_SAMPLE(8); break;
    default: YP_SAMPLE(0); break;
  }
#undef YP_SAMPLE
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_match_partial(const float* d1, const int32_t* n1, int32_t n1_cap, const float* d2, const int32_t* n2, int32_t n2_cap,
                                int32_t D, int32_t col_off, unsigned long long* row_key, unsigned long long* col_key, void* stream) {
  YP_REQUIRE(d1 && d2 && row_key && col_key, YP_ERR_ARG, "match: null pointer");
  YP_REQUIRE(n1_cap > 0 && n2_cap > 0 && D > 0 && D % 4 == 0, YP_ERR_SHAPE, "match: n1=%d n2=%d D=%d (D must be a multiple of 4)", n1_cap, n2_cap, D);
  YP_REQUIRE(yp::aligned16(d1) && yp::aligned16(d2), YP_ERR_ALIGN, "match: descriptors not 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nmax = n1_cap > n2_cap ? n1_cap : n2_cap;
  yp::match_init_kernel<<<yp::ceil_div(nmax, 256), 256, 0, st>>>(row_key, n1_cap, col_key, n2_cap);
  yp::match_tile_kernel<<<dim3(yp::ceil_div(n2_cap, yp::TN), yp::ceil_div(n1_cap, yp::TM)), 256, 0, st>>>(d1, n1, n1_cap, d2, n2, n2_cap, D,
                                                                                                         col_off, row_key, col_key);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_match_finalize(const unsigned long long* row_key, const int32_t* n1, int32_t n1_cap, const unsigned long long* col_key,
                                 int32_t n2_total, float nn_thresh, float* matches, int32_t* match_count, void* stream) {
  YP_REQUIRE(row_key && col_key && matches && match_count, YP_ERR_ARG, "match_finalize: null pointer");
  YP_REQUIRE(nn_thresh >= 0.0f, YP_ERR_ARG, "'nn_thresh' should be non-negative");
  YP_REQUIRE(n1_cap > 0 && n2_total >= 0, YP_ERR_SHAPE, "match_finalize: bad sizes");
  yp::match_finalize_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(row_key, n1, n1_cap, col_key, n2_total, nn_thresh, matches,
                                                                               match_count, 0, 0, nullptr, nullptr, nullptr);
  YP_LAUNCH_OK();
  return YP_OK;
}

namespace {
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  YP_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
  return YP_OK;
}
}  // namespace

extern "C" int yp_match_frames(const float* d1, const int32_t* sel1, const int32_t* n1, const float* d2, const int32_t* sel2, const int32_t* n2,
                               int32_t B, int32_t cap, int32_t D, unsigned long long* row_key, unsigned long long* col_key, float nn_thresh,
                               float* matches, int32_t* match_count, const int32_t* kcount, const int32_t* bcount, int32_t* counts3,
                               void* stream) {
  YP_REQUIRE(d1 && d2 && n1 && n2 && row_key && col_key && matches && match_count, YP_ERR_ARG, "match_frames: null pointer");
  YP_REQUIRE(B > 0 && cap > 0 && D > 0 && D % 4 == 0, YP_ERR_SHAPE, "match_frames: B=%d cap=%d D=%d (D must be a multiple of 4)", B, cap, D);
  YP_REQUIRE(yp::aligned16(d1) && yp::aligned16(d2), YP_ERR_ALIGN, "match_frames: descriptors not 16-byte aligned");
  YP_REQUIRE(nn_thresh >= 0.0f, YP_ERR_ARG, "'nn_thresh' should be non-negative");
  YP_REQUIRE(!counts3 || (kcount && bcount), YP_ERR_ARG, "match_frames: counts3 needs kcount and bcount");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int per_img = std::max(1, 2 * yp::sm_count() / B);
  int rc = launch_pdl(yp::match_frames_kernel, dim3(per_img, B), dim3(256), st, d1, sel1, n1, d2, sel2, n2, cap, D, row_key, col_key);
  if (rc != YP_OK) return rc;
  return launch_pdl(yp::match_finalize_kernel, dim3(B), dim3(1024), st, static_cast<const unsigned long long*>(row_key), n1, cap,
                    static_cast<const unsigned long long*>(col_key), cap, nn_thresh, matches, match_count, static_cast<long long>(cap),
                    static_cast<long long>(cap) * 3, kcount, bcount, counts3);
}

extern "C" int yp_gather_rows(const float* src, const int32_t* sel, const int32_t* count, int32_t B, int32_t cap, int32_t D, float* dst, void* stream) {
  YP_REQUIRE(src && sel && count && dst, YP_ERR_ARG, "gather_rows: null pointer");
  YP_REQUIRE(B > 0 && cap > 0 && D > 0 && D % 4 == 0, YP_ERR_SHAPE, "gather_rows: B=%d cap=%d D=%d", B, cap, D);
  YP_REQUIRE(yp::aligned16(src) && yp::aligned16(dst), YP_ERR_ALIGN, "gather_rows: buffers not 16-byte aligned");
  yp::gather_rows_kernel<<<dim3(std::max(1, 2 * yp::sm_count() / B), B), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, sel, count, cap, D, dst);
  YP_LAUNCH_OK();
  return YP_OK;
}
