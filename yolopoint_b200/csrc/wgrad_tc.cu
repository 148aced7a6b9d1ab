// tcgen05 weight gradient of a convolution for sm_100a  (yp_conv2d_nhwc_wgrad).
//
//   dW[co][tap][ci] += sum over (b, oh, ow) of dY[b, oh, ow, co] * X[b, oh*s + kh - p, ow*s + kw - p, ci]
//
// GEMM view: M = co (tile 128), N = ci (tile 64 or 128), K = output pixels.  Both operands are NHWC, i.e. the K index (pixel) is
// the slow dimension and the channels are contiguous: they are "MN-major" UMMA operands.  A TMA box (64 channels, Wp, Ht) lands
// in shared memory as [pixel rows][64 bf16 = 128 bytes] with the 128-byte swizzle, which is exactly the canonical MN-major
// SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO)) (cute/atom/mma_traits_sm100.hpp): SBO = 1024 bytes (8 pixel rows), LBO =
// the distance between two 64-channel blocks.
//
// K tiles are row strips: Ht output rows x Wp columns, Wp = Wo + 2 for 3x3 filters; the extra columns are out of bounds for dY
// and therefore zero-filled by TMA, so that pixel p = h * Wp + w of the dY strip pairs with row p + kw of the X strip that was
// loaded one column to the left: the three taps of a filter row are row-shifted windows of ONE X strip (the tensor core derives
// the swizzle phase from the absolute shared-memory address, see conv_tc.cu patch mode).  Filter rows (kh) are spread over
// grid.z; stride 2 reads the even/odd parity views of X (two strips: even columns for kw = 1, odd columns for kw = 0 / 2).
//
// The pixel range is split over grid.x; every CTA adds its fp32 partial tile to dW with vector reductions (red.global.add.v4.f32).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>

#include <mutex>
#include <stdlib.h>

namespace yp {
namespace {

struct alignas(64) WgradMaps {
  CUtensorMap dy;      // (C, W, H, B) of dY
  CUtensorMap x[4];    // X strips: stride 1: [0]; stride 2: [row parity * 2 + column parity] views of the input
};

struct WgradArgs {
  int Ho, Wo, Wp, Ht, strips_per_img, n_strips, strips_per_cta;
  int Cin, Cout, Ktot;             // Ktot = taps * Cin (row length of dW)
  int ksize, stride, kh0;
  int NB;                          // 64-channel ci blocks per CTA (N = 64 * NB)
  int n_box;                       // X strips per stage (1, or 2 for 3x3 stride 2)
  int n_taps;                      // taps per CTA (1 or 3)
  int tap_box[3], tap_shift[3];    // per tap: X strip and row shift inside it
  int x_w0[2];                     // W coordinate the X strips start at (-1 or 0)
  int Kp;                          // K rows per strip, padded to 16
  int dy_blk_bytes, x_blk_bytes;   // bytes of one 64-channel block of the dY / X strip (LBO)
  int stage_bytes, stages, bar_off, tx_bytes;
  uint32_t tmem_cols, idesc;
  float* dw;
};

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 4;

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradMaps maps, const WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp index as a warp-uniform value

  const int n_ci_tiles = (a.Cin + 64 * a.NB - 1) / (64 * a.NB);
  const int co0 = (blockIdx.y / n_ci_tiles) * 128;
  const int ci0 = (blockIdx.y % n_ci_tiles) * 64 * a.NB;
  const int kh = a.kh0 + blockIdx.z;
  const int strip0 = blockIdx.x * a.strips_per_cta;
  const int n_my = min(a.n_strips, strip0 + a.strips_per_cta) - strip0;

  const uint32_t bar_base = smem_base + a.bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWgMaxStages + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * kWgMaxStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kWgMaxStages + 1);

  // Rows [Ht*Wp, Kp) of every block are never written by TMA: they must read as zero (dY) / finite (X).
  {
    uint4* p = reinterpret_cast<uint4*>(smem_gen);
    const int n16 = a.stages * a.stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += kWgThreads) p[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.dy);
    for (int i = 0; i < (a.stride == 2 ? 4 : 1); ++i) tma_prefetch_desc(&maps.x[i]);
    for (int s = 0; s < kWgMaxStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
  fence_proxy_async_smem();   // the zero fill (generic proxy) is ordered before the TMA writes (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // vertical placement of the X strip for this filter row
  int x_dh = 0;
  if (a.ksize == 3) x_dh = a.stride == 1 ? kh - 1 : (kh == 0 ? -1 : 0);
  const uint32_t x_off = 2u * a.dy_blk_bytes;                 // X strips follow the two dY blocks inside a stage
  const uint32_t x_box_bytes = a.NB * a.x_blk_bytes;

  // Producer and issuer loops are warp-uniform (the whole warp walks the loop, one elected lane executes each TMA / tcgen05
  // instruction, 32-bit descriptor arithmetic) -- see conv_tc.cu: issuing under `lane == 0` costs ~450 cycles per MMA.
  if (warp == 0) {
    // ===================== TMA producer =====================
    int s = 0, ph = 0;
    for (int i = 0; i < n_my; ++i) {
      const int strip = strip0 + i;
      const int b = strip / a.strips_per_img;
      const int h0 = (strip - b * a.strips_per_img) * a.Ht;
      mbar_wait(empty_bar(s), ph ^ 1);
      const uint32_t st = smem_base + s * a.stage_bytes;
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), a.tx_bytes);
        for (int j = 0; j < 2; ++j) tma_load_4d(st + j * a.dy_blk_bytes, &maps.dy, full_bar(s), co0 + 64 * j, 0, h0, b);
        for (int bx = 0; bx < a.n_box; ++bx) {
          // stride 2: filter rows kh = 0 / 2 read the odd input rows, kh = 1 the even ones; strip bx = column parity
          const CUtensorMap* m = &maps.x[(a.stride == 2 ? (kh == 1 ? 0 : 2) : 0) + bx];
          for (int j = 0; j < a.NB; ++j)
            tma_load_4d(st + x_off + bx * x_box_bytes + j * a.x_blk_bytes, m, full_bar(s), ci0 + 64 * j, a.x_w0[bx], h0 + x_dh, b);
        }
      }
      if (++s == a.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const int ksteps = a.Kp / 16;
    const uint32_t ncol = 64u * a.NB;
    // MN-major SWIZZLE_128B descriptor halves: low = start >> 4 | LBO >> 4 << 16, high = SBO (1024 B) | version 1 | layout 2
    const uint32_t dhi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lbo = static_cast<uint32_t>(a.dy_blk_bytes >> 4) << 16, b_lbo = static_cast<uint32_t>(a.x_blk_bytes >> 4) << 16;
    int s = 0, ph = 0;
    for (int i = 0; i < n_my; ++i) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t st = smem_base + s * a.stage_bytes;
      for (int t = 0; t < a.n_taps; ++t) {
        const uint32_t xb = st + x_off + a.tap_box[t] * x_box_bytes + a.tap_shift[t] * 128u;
        uint32_t al = ((st & 0x3FFFFu) >> 4) | a_lbo, bl = ((xb & 0x3FFFFu) >> 4) | b_lbo;
        const uint32_t dcol = tmem_base + t * ncol;
        for (int k = 0; k < ksteps; ++k) {
          umma32<false>(dcol, al, bl, dhi, a.idesc, (i | k) ? 1u : 0u);
          al += (16u * 128u) >> 4; bl += (16u * 128u) >> 4;   // 16 pixel rows of 128 bytes
        }
      }
      umma_commit_elect(empty_bar(s));
      if (++s == a.stages) { s = 0; ph ^= 1; }
    }
    umma_commit_elect(accum_bar);
  }
  __syncwarp();
  if (warp >= 2) {
    // ===================== epilogue: TMEM -> red.global.add =====================
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int ncol = 64 * a.NB;
    for (int t = 0; t < a.n_taps; ++t) {
      const int tap = a.ksize == 3 ? kh * 3 + t : 0;
      float* row = a.dw + static_cast<long long>(co) * a.Ktot + static_cast<long long>(tap) * a.Cin + ci0;
      for (int c = 0; c < ncol; c += 16) {
        float v[16];
        tmem_ld16(taddr_row + t * ncol + c, v);   // warp-collective: executed by all lanes
        tmem_ld_wait();
        if (co < a.Cout && ci0 + c < a.Cin) {
#pragma unroll
          for (int j = 0; j < 4; ++j) red_add_v4(row + c + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

PFN_cuTensorMapEncodeTiled_v12000 g_wg_encode = nullptr;
std::once_flag g_wg_once;
PFN_cuTensorMapEncodeTiled_v12000 wg_encode() {
  std::call_once(g_wg_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      g_wg_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  return g_wg_encode;
}

// 4-D map (C, W, H, B) over a bf16 NHWC view.  sub = 2 selects the (ph, pw) parity view; `col0` / `n_cols` restrict the map to
// the columns [col0, col0 + n_cols) of that (sub-sampled) view, so that everything outside reads as zero (TMA out-of-bounds fill).
int encode_nhwc(CUtensorMap* tm, const YpView& v, int sub, int ph, int pw, int col0, int n_cols, int box_w, int box_h) {
  char* base = static_cast<char*>(v.base) + ((static_cast<int64_t>(ph) * v.W + pw) + static_cast<int64_t>(col0) * sub) * v.pix_stride * 2;
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)n_cols, (cuuint64_t)(v.H / sub), (cuuint64_t)v.B};
  cuuint64_t strides[3] = {(cuuint64_t)(v.pix_stride * sub * 2), (cuuint64_t)(v.pix_stride * v.W * sub * 2),
                           (cuuint64_t)(static_cast<int64_t>(v.H) * v.W * v.pix_stride * 2)};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  YP_REQUIRE(aligned16(base), YP_ERR_ALIGN, "wgrad: view base %p not 16-byte aligned", (void*)base);
  for (int i = 0; i < 3; ++i) YP_REQUIRE(strides[i] % 16 == 0, YP_ERR_ALIGN, "wgrad: view stride %d (%llu B) not a multiple of 16", i, (unsigned long long)strides[i]);
  CUresult r = wg_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YP_REQUIRE(r == CUDA_SUCCESS, YP_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad) failed: %d (C=%d cols=%d H=%d B=%d box %d,%d)", (int)r, v.C, n_cols,
             v.H / sub, v.B, box_w, box_h);
  return YP_OK;
}

// One launch (three for stride 2: one per filter row) over the output columns [c0, c1) of every row.
int wgrad_segment(const YpWgradDesc& d, int c0, int c1, cudaStream_t st) {
  const YpView& x = d.x;
  const YpView& dy = d.dy;
  WgradArgs a;
  memset(&a, 0, sizeof(a));
  const int Wfull = dy.W;                       // output columns of the whole layer
  a.Ho = dy.H; a.Wo = c1 - c0; a.Cin = x.C; a.Cout = dy.C; a.ksize = d.ksize; a.stride = d.stride;
  a.Ktot = d.ksize * d.ksize * x.C;
  a.dw = d.dw;
  a.Wp = d.ksize == 3 ? a.Wo + 2 : a.Wo;
  // X strips.  Stride 1 (3x3): one strip whose row j holds input column c0 - 1 + j.  Stride 2: strip 0 = even input columns from
  // output column c0 (tap kw = 1), strip 1 = odd input columns from c0 - 1 (kw = 0 -> shift 0, kw = 2 -> shift 1).  A strip that
  // would start left of the image starts at the image edge and is loaded at coordinate -1 instead (zero fill = conv padding).
  int x_col0[2] = {c0, c0}, x_sub = d.stride;
  if (d.ksize == 1) {
    a.n_taps = 1; a.n_box = 1; a.tap_box[0] = 0; a.tap_shift[0] = 0; a.x_w0[0] = 0;
  } else if (d.stride == 1) {
    a.n_taps = 3; a.n_box = 1;
    for (int t = 0; t < 3; ++t) { a.tap_box[t] = 0; a.tap_shift[t] = t; }
    if (c0 > 0) { x_col0[0] = c0 - 1; a.x_w0[0] = 0; } else { x_col0[0] = 0; a.x_w0[0] = -1; }
  } else {
    a.n_taps = 3; a.n_box = 2; a.x_w0[0] = 0;
    a.tap_box[0] = 1; a.tap_shift[0] = 0;
    a.tap_box[1] = 0; a.tap_shift[1] = 0;
    a.tap_box[2] = 1; a.tap_shift[2] = 1;
    if (c0 > 0) { x_col0[1] = c0 - 1; a.x_w0[1] = 0; } else { x_col0[1] = 0; a.x_w0[1] = -1; }
  }
  // tile geometry: the widest N (ci) and tallest strip that leave >= 2 pipeline stages
  const int budget = 200 * 1024;
  int best_NB = 0, best_Ht = 0, best_stages = 0, best_Kp = 0;
  for (int NB = (x.C > 64 ? 2 : 1); NB >= 1 && !best_NB; --NB) {
    for (int Ht = std::min(a.Ho, 256); Ht >= 1; --Ht) {
      const int Kp = (Ht * a.Wp + 15) & ~15;
      const int stage = 2 * Kp * 128 + a.n_box * NB * (Kp + 8) * 128;
      if (Kp > 256 && Ht > 1) continue;                 // long strips only lengthen the pipeline fill
      const int stages = std::min(budget / stage, kWgMaxStages);
      if (stages >= (Kp <= 128 ? 3 : 2)) { best_NB = NB; best_Ht = Ht; best_stages = stages; best_Kp = Kp; break; }
    }
  }
  YP_REQUIRE(best_NB > 0, YP_ERR_SHAPE, "wgrad: no strip geometry fits shared memory (segment of %d columns)", a.Wo);
  a.NB = best_NB; a.Ht = best_Ht; a.stages = best_stages; a.Kp = best_Kp;
  a.dy_blk_bytes = a.Kp * 128;
  a.x_blk_bytes = (a.Kp + 8) * 128;
  a.stage_bytes = 2 * a.dy_blk_bytes + a.n_box * a.NB * a.x_blk_bytes;
  a.tx_bytes = (2 + a.n_box * a.NB) * a.Ht * a.Wp * 128;   // TMA counts the full box, out-of-bounds elements included
  a.bar_off = a.stages * a.stage_bytes;
  const size_t smem = 1024 + a.bar_off + 8 * (2 * kWgMaxStages + 2) + 16;
  YP_REQUIRE(smem <= 227 * 1024, YP_ERR_SHAPE, "wgrad: needs %zu bytes of shared memory", smem);
  const int cols = a.n_taps * 64 * a.NB;
  a.tmem_cols = 32;
  while ((int)a.tmem_cols < cols) a.tmem_cols <<= 1;
  // instruction descriptor: D fp32, A/B bf16, both MN-major (bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>((64 * a.NB) >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);

  a.strips_per_img = ceil_div(a.Ho, a.Ht);
  a.n_strips = a.strips_per_img * x.B;
  const int tiles_y = ceil_div(a.Cout, 128) * ceil_div(a.Cin, 64 * a.NB);
  const int tiles_z = d.ksize == 3 ? 3 : 1;
  // One CTA per SM is resident (3 x ~50 KB stages): size the pixel split so that the grid is a whole number of waves -- rounding
  // up (e.g. 99 x 3 = 297 CTAs on 148 SMs) would add a third, nearly empty wave.
  // A single wave: every CTA pays one epilogue (128 x 384 fp32 reductions into dW); measured on the YOLOPoint-L layers, one wave of
  // long CTAs beats two or three waves (3.20 vs 3.88 vs 4.71 ms per backward pass).
  static const int waves = getenv("YP_WGRAD_WAVES") ? atoi(getenv("YP_WGRAD_WAVES")) : 1;
  int P = (waves * sm_count()) / (tiles_y * tiles_z);
  if (P > a.n_strips) P = a.n_strips;
  if (P < 1) P = 1;
  a.strips_per_cta = ceil_div(a.n_strips, P);
  P = ceil_div(a.n_strips, a.strips_per_cta);

  WgradMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  if ((rc = encode_nhwc(&maps.dy, dy, 1, 0, 0, c0, a.Wo, a.Wp, a.Ht)) != YP_OK) return rc;
  const int Wx = x.W / x_sub;                   // columns of the (sub-sampled) input view
  static thread_local bool configured = false;
  if (!configured) {
    YP_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  (void)Wfull;
  if (d.stride == 1) {
    if ((rc = encode_nhwc(&maps.x[0], x, 1, 0, 0, x_col0[0], Wx - x_col0[0], a.Wp, a.Ht)) != YP_OK) return rc;
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw)
        if ((rc = encode_nhwc(&maps.x[ph * 2 + pw], x, 2, ph, pw, x_col0[pw], Wx - x_col0[pw], a.Wp, a.Ht)) != YP_OK) return rc;
  }
  wgrad_tc_kernel<<<dim3(P, tiles_y, tiles_z), kWgThreads, smem, st>>>(maps, a);
  YP_LAUNCH_OK();
  return YP_OK;
}

}  // namespace

int wgrad_tc(const YpWgradDesc& d, cudaStream_t st) {
  YP_REQUIRE(wg_encode() != nullptr, YP_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const YpView& x = d.x;
  const YpView& dy = d.dy;
  YP_REQUIRE(x.format == YP_FMT_BF16 && dy.format == YP_FMT_BF16, YP_ERR_SHAPE, "wgrad: operands must be bf16");
  YP_REQUIRE((d.ksize == 1 && d.stride == 1) || (d.ksize == 3 && (d.stride == 1 || d.stride == 2)), YP_ERR_SHAPE, "wgrad: k=%d s=%d unsupported",
             d.ksize, d.stride);
  YP_REQUIRE(x.C % 8 == 0 && dy.C % 8 == 0, YP_ERR_SHAPE, "wgrad: Cin=%d / Cout=%d must be multiples of 8", x.C, dy.C);
  YP_REQUIRE(d.stride == 1 || (x.H % 2 == 0 && x.W % 2 == 0), YP_ERR_SHAPE, "wgrad: stride 2 needs even H,W");
  YP_REQUIRE(dy.B == x.B && dy.H == x.H / d.stride && dy.W == x.W / d.stride, YP_ERR_SHAPE, "wgrad: dY geometry %dx%dx%d does not match X %dx%dx%d / s%d",
             dy.B, dy.H, dy.W, x.B, x.H, x.W, d.stride);
  YP_REQUIRE(aligned16(d.dw), YP_ERR_ALIGN, "wgrad: dW not 16-byte aligned");
  // A strip spans whole output rows (the zero columns right of the row are what lets one X strip serve three taps); rows wider
  // than one TMA box (256 columns) are processed as column segments, each with its own launch.
  const int max_cols = d.ksize == 3 ? 254 : 256;
  const int n_seg = ceil_div(dy.W, max_cols);
  const int seg = ceil_div(dy.W, n_seg);
  for (int c0 = 0; c0 < dy.W; c0 += seg) {
    const int rc = wgrad_segment(d, c0, std::min(dy.W, c0 + seg), st);
    if (rc != YP_OK) return rc;
  }
  return YP_OK;
}

}  // namespace yp
