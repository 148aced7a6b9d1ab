// Descriptor sampling (bilinear grid_sample + L2 normalise) and brute-force two-way matching
// (demo.py:200-215, 300-341; evaluations/descriptor_evaluation.py:148-181 of the reference).
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
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.0f;
  if (threadIdx.x < TM) rmin[threadIdx.x] = ~0ull;
  if (threadIdx.x < TN) cmin[threadIdx.x] = ~0ull;
  for (int k0 = 0; k0 < D; k0 += TK) {
    // 64 rows x 16 k: each thread loads one float4 of A and one of B (rows are D-contiguous)
    {
      const int r = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (i0 + r < n1 && k0 + kq < D) va = *reinterpret_cast<const float4*>(d1 + static_cast<int64_t>(i0 + r) * D + k0 + kq);
      if (j0 + r < n2 && k0 + kq < D) vb = *reinterpret_cast<const float4*>(d2 + static_cast<int64_t>(j0 + r) * D + k0 + kq);
      As[kq][r] = va.x; As[kq + 1][r] = va.y; As[kq + 2][r] = va.z; As[kq + 3][r] = va.w;
      Bs[kq][r] = vb.x; Bs[kq + 1][r] = vb.y; Bs[kq + 2][r] = vb.z; Bs[kq + 3][r] = vb.w;
    }
    __syncthreads();
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
}

// ascending-i compaction of mutual nearest neighbours below the threshold (single block)
__global__ void match_finalize_kernel(const unsigned long long* __restrict__ row_key, const int* __restrict__ n1p, int n1_cap,
                                      const unsigned long long* __restrict__ col_key, int n2_total, float thr,
                                      float* __restrict__ matches, int* __restrict__ match_count) {
  __shared__ int warp_excl[32];
  __shared__ int block_total;
  const int n1 = n1p ? min(*n1p, n1_cap) : n1_cap;
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
  if (threadIdx.x == 0) *match_count = carry;
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
                                                                               match_count);
  YP_LAUNCH_OK();
  return YP_OK;
}
