// tcgen05 implicit-GEMM convolution for sm_100a  (yp_conv2d_nhwc_fwd, algo YP_ALGO_TCGEN05).
//
// GEMM view: M = output pixels (one CTA = one Ht x Wt patch of one image, <= 128 pixels, the UMMA M=128
// tile), N = output channels (tile Nt <= 256), K = taps * Cin.  The A operand is never materialised:
// for every filter tap a TMA *tiled* load of the box (Ck channels, Wt, Ht) at the tap-shifted coordinate
// lands in shared memory in exactly the K-major 128B/64B/32B-swizzled layout tcgen05.mma consumes; the
// conv zero padding is TMA out-of-bounds fill.  Stride-2 convs read four parity views of the input
// (even/odd rows x even/odd columns), each a plain strided 5-D tensor map, so no im2col mode is needed.
//
// Numerics: YP_FMT_F32X2 operands are (hi, lo) TF32 pairs; the kernel issues A_lo*W_hi + A_hi*W_lo +
// A_hi*W_hi into one fp32 TMEM accumulator ("3xTF32", ~fp32 accuracy); YP_FMT_BF16 issues one bf16 MMA.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warp 2 = second MMA issuer; when their
// loops are done all eight warps run the epilogue (TMEM -> registers -> bias/SiLU/residual/L2-norm -> swizzled smem -> TMA store
// to 1..8 maps: channel-slice (concat) destinations and the four parity views of a 2x-upsampled destination), two warps per
// TMEM lane quarter.  conv_tc_persist_kernel (bf16 layers with several waves of tiles) keeps producer / issuers / epilogue warps
// separate and loops over tiles with double-buffered TMEM accumulators.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "sppf.cuh"

#include <cudaTypedefs.h>

#include <algorithm>
#include <mutex>
#include <type_traits>
#include <vector>
#include <stdlib.h>

namespace yp {
namespace {

// ---------------------------------------------------------------------------------------------
// Kernel arguments
// ---------------------------------------------------------------------------------------------
struct alignas(64) ConvMaps {
  CUtensorMap in[4];   // [0] for stride 1; (ph*2+pw) parity views for stride 2
  CUtensorMap w;       // packed weights (K, Cout, plane)
  CUtensorMap out[8];  // destinations: per output either 1 map or 4 parity maps (2x upsample)
};

struct ConvArgs {
  int tiles_w, tiles_h, Ht, Wt, Ho, Wo;
  int ksize, stride;
  unsigned long long tap_dh, tap_dw;   // custom tap list (ksize == 0): 4 bits per tap, offset + 8
  int n_taps, kb_per_tap, ck_bytes, ck_elems;
  int in_planes, a_split;          // a_split: 1 = both planes of A arrive with one TMA (box plane dim 2)
  int a_plane_off;                 // smem byte offset of the lo plane inside the A region
  int Nt, stages, stage_bytes, a_region_bytes, bar_off, tx_bytes;
  // MMA issuers (see kernel comment).  Issuer q emits, per 32-byte k-step, n_jobs[q] MMAs (A plane ja, W plane jb) with
  // instruction descriptor iss_idesc[q] into accumulator (iss_col[q] + r * iss_stride[q]), r rotating over iss_cnt[q].
  // Issuer q (warp 1 + q) takes the k-steps kstart[q], kstart[q] + kinc[q], ... of every k-block.
  int n_iss;
  int kstart[4], kinc[4];
  int n_jobs[4], job_a[4][2], job_b[4][2];
  int iss_col[4], iss_stride[4], iss_cnt[4];
  uint32_t iss_idesc[4];
  int n_src, src_col[16];          // TMEM column bases whose sum is the result (main products first)
  int split_k, kb_per_split;       // grid.z CTAs share one output tile, each reducing a slice of K
  // patch mode (3x3 stride 1): one (Ht+2) x (Wt+2) input patch per channel block stays in shared memory and the nine taps
  // are row-shifted windows of it; M rows index the padded patch (row = h * Wp + w), halo columns are discarded.
  int patch, Wp, a_rows, a_stage_bytes, a_tx, b_stage_bytes, b_tx, b_stages, b_ring_off, base_offset_mode;
  float* ws_partial;               // [split][m_tile][n_tile][128][Nt] fp32 partial sums
  int* ws_counter;                 // [m_tile][n_tile] arrival counters (self-resetting)
  uint32_t tmem_cols;
  const float* bias;
  int act, l2norm;
  int rowmin, col_off;                 // YP_EPI_ROWMIN: reduce distance keys over the output channels instead of storing
  unsigned long long* row_key;
  unsigned long long* col_key;         // optional: column minima of the same tile (the other direction of the two-way match)
  const int* n_rows;
  const int* n_cols;
  const void* res_base;
  long long res_pix, res_plane;
  int n_out_maps, out_planes, out_row_bytes, staging_set_bytes, out_fmt;
  int drain_units;   // conv_tc_drain_kernel (persist == 2): pipeline units per accumulation round
  int persist, n_tiles_n, tiles_total, stg_off, buf_cols;   // persistent variant: tiles per CTA loop, staging offset, TMEM columns per accumulator buffer
  int ablate;      // debug (YP_CONV_ABLATE): 1 = issue no MMAs, 2 = issue no TMA loads (timing experiments; results are garbage)
  long long* dbg;  // optional timeline buffer (yp_debug_conv_timeline); CTA (0,0) records clock64 stamps
};

// CTA size: 256 threads (8 warps: warp 0 TMA producer, warps 1-2 MMA issuers, then all run the epilogue; two CTAs per SM) or 512 (one CTA per SM)
constexpr size_t kWsCounterBytes = 64 * 1024;   // split-K arrival counters: one int per output tile, <= 16384 tiles
constexpr int kMaxStages = 8;
constexpr int kPersistThreads = 384;   // persistent variant: warp 0 producer, warps 1-2 issuers, warps 4-11 epilogue

template <int OUT_FMT>
struct OutT { using type = float; };
template <>
struct OutT<YP_FMT_BF16> { using type = __nv_bfloat16; };

// branch-free SiLU with ~1 ulp sigmoid: ex2.approx (rel err 2^-22) + rcp.approx refined by one Newton step
__device__ __forceinline__ float silu_fast(float v) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-1.4426950408889634f * v));
  const float d = 1.0f + t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = fmaf(r, fmaf(-d, r, 1.0f), r);
  return v * r;
}

// ---------------------------------------------------------------------------------------------
// Kernel.  UNITS = 16-column TMEM units per staging row chunk (chunk_elems = 16 * UNITS).
//
// Accumulators.  A tcgen05.mma that accumulates into the tile written by the previous MMA waits for it
// (~140 cycles measured), far longer than the math of one small MMA, and the tensor core adds into an fp32
// accumulator with truncation, so the error grows linearly with the number of MMAs chained on one accumulator.
// Both are cured by spreading the MMAs of a tile over several independent TMEM accumulators that the epilogue
// sums with round-to-nearest adds: `n_main` accumulators take the main products round-robin; in 3xTF32 mode
// `n_small` (<= 2) more take the two cross terms (A_lo*W_hi, A_hi*W_lo), whose partial sums are ~2^-11 of the
// result so that their truncation error is negligible.
// ---------------------------------------------------------------------------------------------
// Chain mode reads the layer's arguments from an element of an array in the kernel-parameter space (dynamic constant-bank index):
// the compiler then re-loads loop-invariant fields inside the single-thread producer / issuer loops (LDC / LDCU + R2UR per
// k-block) where the stand-alone kernel has immediate constant operands.  `held` pins such a value in a register before the loop.
template <bool kHold>
__device__ __forceinline__ uint32_t held(uint32_t v) {
  if (kHold) asm volatile("" : "+r"(v));
  return v;
}

// One work item = what one CTA of the stand-alone launch does (bx, by, bz = its blockIdx; gx, gy = the launch's grid.x / grid.y).
// kChain: the item runs inside conv_chain_kernel (a persistent CTA walking the layers of a network segment): TMEM is allocated once
// per CTA (chain_tmem), the mbarriers are re-initialised per item, there is no programmatic dependent launch, `dep_wait` blocks until
// the layers this one reads from have completed, and the item ends with all of its global writes complete (the caller publishes the
// layer's completion counter).
template <int OUT_FMT, int UNITS, bool kTf32, int NT, bool kChain, class WaitFn>
__device__ __forceinline__ void conv_tile(const ConvMaps& maps, const ConvArgs& a, const int bx, const int by, const int bz, const int gx, const int gy,
                                          uint8_t* smem_raw, const uint32_t chain_tmem, const bool chain_first, WaitFn&& dep_wait,
                                          long long* chain_dbg = nullptr, const int chain_variant = 0) {
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp index as a warp-uniform value
  const int per_img = a.tiles_w * a.tiles_h;
  const int b = bx / per_img;
  const int trem = bx - b * per_img;
  const int th = trem / a.tiles_w, tw = trem - th * a.tiles_w;
  const int h0 = th * a.Ht, w0 = tw * a.Wt;
  const int n0 = by * a.Nt;
  const int num_kb_all = a.patch ? a.kb_per_tap : a.n_taps * a.kb_per_tap;   // K-loop units (patch mode: channel blocks)
  const int kb0 = bz * a.kb_per_split;                       // this CTA's slice of the K loop (split-K)
  const int num_kb = min(num_kb_all, kb0 + a.kb_per_split) - kb0;
  long long* dbg = kChain ? chain_dbg : ((a.dbg && bx == 0 && by == 0 && bz == 0) ? a.dbg : nullptr);
  auto stamp = [&](int slot) {   // callers restrict it to one lane; chain mode: 8 slots per item, nanoseconds
    if (!dbg) return;
    if (kChain) {   // 8 first TMA issued, 9 first operands landed, 10 first k-block issued, 11 all MMAs issued
      const int idx = slot < 6 ? slot : (slot == 8 ? 8 : (slot == 104 ? 9 : (slot == 200 ? 10 : (slot == 400 ? 11 : -1))));
      if (idx >= 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[idx] = static_cast<long long>(t); }
    } else {
      dbg[slot] = clock64();
    }
  };
  if (threadIdx.x == 0) stamp(0);
  // whole-grid occupancy picture: every CTA records (SM id, start, end) in nanoseconds at dbg[512 + 3 * linear CTA index]
  const long long cta_lin = (static_cast<long long>(bz) * gy + by) * gx + bx;
  if (!kChain && a.dbg && threadIdx.x == 0 && cta_lin < 20000) {
    unsigned smid; unsigned long long t;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg[512 + 3 * cta_lin] = smid; a.dbg[513 + 3 * cta_lin] = static_cast<long long>(t);
  }
  // Programmatic dependent launch: let the next kernel of the stream start its prologue (barrier init, TMEM allocation,
  // descriptor prefetch) now; it blocks in griddepcontrol.wait until this grid has completed and flushed.
  if (!kChain) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // barriers + tmem slot + bias live after the pipeline/staging region
  const uint32_t bar_base = smem_base + a.bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * kMaxStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 1);
  volatile int* split_flag = reinterpret_cast<volatile int*>(smem_gen + a.bar_off + 8 * (2 * kMaxStages + 1) + 4);
  auto afull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 4 + s); };
  float* bias_s = reinterpret_cast<float*>(smem_gen + a.bar_off + 8 * (2 * kMaxStages + 6));

  if (warp == 0 && lane == 0) {
    const int n_in = a.stride == 2 ? 4 : 1;
    for (int i = 0; i < n_in; ++i) tma_prefetch_desc(&maps.in[i]);
    tma_prefetch_desc(&maps.w);
    for (int i = 0; i < a.n_out_maps; ++i) tma_prefetch_desc(&maps.out[i]);
    if (kChain && !chain_first) {   // the previous item's barriers: every arrival on them has happened (see the end of the item)
      for (int s = 0; s < 2 * kMaxStages + 1; ++s) mbar_inval(bar_base + 8u * s);
      for (int s = 0; s < 4; ++s) mbar_inval(bar_base + 8u * (2 * kMaxStages + 2 + s));
    }
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), a.n_iss); }
    for (int s = 0; s < 2; ++s) { mbar_init(afull_bar(s), 1); mbar_init(aempty_bar(s), a.n_iss); }
    mbar_init(accum_bar, a.n_iss);
    fence_barrier_init();
  }
  if (!kChain && warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
  if (kChain) dep_wait();   // one thread spins on the completion counters of the producer layers (under the barrier set-up of the others)
  for (int i = threadIdx.x; i < a.Nt; i += NT) bias_s[i] = a.bias ? a.bias[n0 + i] : 0.0f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base = chain_tmem;
  if (!kChain) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);   // same value in every lane: lets the compiler keep MMA operands in uniform registers
  if (threadIdx.x == 0) stamp(1);
  if (a.ablate != 6) {   // 6 = launch + prologue only (timing experiment)

  // ---- epilogue geometry of this thread (needed before the roles start: the residual of the thread's first unit is fetched NOW,
  // under the main loop -- it cost 0.4-0.8 us of exposed latency in front of the first epilogue barrier, longest for the producer /
  // issuer warps, which only reach the epilogue when their loops are done).  Unit u = columns [16 u, 16 u + 16); first unit = hf.
  constexpr int NG = NT / 128;                   // warp groups (2 or 4)
  const int q = warp & 3;                        // TMEM lane quarter this warp may access
  const int hf = warp >> 2;                      // warp group
  const int row = q * 32 + lane;               // accumulator row (TMEM lane) of this thread
  const int rdiv = a.patch ? a.Wp : a.Wt;      // patch mode: rows index the padded patch, halo columns are dropped
  const int rh = row / rdiv, rw = row - rh * rdiv;
  const int oh = h0 + rh, ow = w0 + rw;
  const bool in_tile = rh < a.Ht && rw < a.Wt;
  const bool valid = in_tile && (oh < a.Ho) && (ow < a.Wo);
  const int srow = in_tile ? rh * a.Wt + rw : 127;   // row of the (compact Ht x Wt) staging tile this thread fills
  const bool et0 = (threadIdx.x == 64);
  const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  const bool has_res = a.res_base != nullptr && valid;
  const long long res_off = ((static_cast<long long>(b) * a.Ho + oh) * a.Wo + ow) * a.res_pix + n0;

  // A chunk (one staging row, CH columns) is computed in passes of 16 columns by a rolled loop: few live registers (two
  // CTAs per SM, so that one CTA's epilogue overlaps the other's main loop) and a loop body that stays in the instruction
  // cache -- the unrolled epilogue ran once per CTA from cold code and spent ~40 % of its issue slots waiting for
  // instruction fetch (ncu source view, stall_no_inst).
  constexpr int PU = 1;                          // 16-column units per pass
  constexpr int PE = 16 * PU;                    // columns per pass
  // residual of columns [col0, col0 + PE) for this thread's pixel, planes summed (residual format == output format family `fmt`:
  // a compile-time constant in the stand-alone kernel, the layer's a.out_fmt for the early fetch of a chained item)
  auto load_res = [&](int fmt, int col0, float* r) {
    if (!has_res) {
#pragma unroll
      for (int i = 0; i < PE; ++i) r[i] = 0.0f;
      return;
    }
    if (fmt == YP_FMT_BF16) {
      const uint4* p = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(a.res_base) + res_off + col0);
#pragma unroll
      for (int j = 0; j < PE / 8; ++j) {
        const uint4 u = __ldg(p + j);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); r[j * 8 + 2 * e] = f.x; r[j * 8 + 2 * e + 1] = f.y; }
      }
    } else {
      const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(a.res_base) + res_off + col0);
      const float4* pl = reinterpret_cast<const float4*>(static_cast<const float*>(a.res_base) + res_off + a.res_plane + col0);
#pragma unroll
      for (int j = 0; j < PE / 4; ++j) {
        const float4 hi = __ldg(p + j);
        float4 lo = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fmt == YP_FMT_F32X2) lo = __ldg(pl + j);
        r[j * 4] = hi.x + lo.x; r[j * 4 + 1] = hi.y + lo.y; r[j * 4 + 2] = hi.z + lo.z; r[j * 4 + 3] = hi.w + lo.w;
      }
    }
  };
  float res[PE];
  const bool res_pre = hf * 16 < a.Nt && a.split_k == 1 && !a.rowmin && !a.l2norm;
  if (res_pre) {
    if (!kChain && a.res_base) asm volatile("griddepcontrol.wait;" ::: "memory");   // the residual may be the previous kernel's output
    load_res(kChain ? a.out_fmt : OUT_FMT, hf * 16, res);
  }

  if (warp == 0 && elect_one()) {
    // ===================== TMA producer (one elected lane runs the whole loop) =====================
    if (!kChain) asm volatile("griddepcontrol.wait;" ::: "memory");   // inputs are written by the previous kernel(s) of the stream
    const bool load = a.ablate != 2;
    const uint32_t p_a_stage = held<kChain>(a.a_stage_bytes), p_a_tx = held<kChain>(a.a_tx), p_ck = held<kChain>(a.ck_elems), p_b_tx = held<kChain>(a.b_tx);
    const uint32_t p_b_ring = held<kChain>(a.b_ring_off), p_b_stage = held<kChain>(a.b_stage_bytes), p_kbt = held<kChain>(a.kb_per_tap);
    const uint32_t p_b_stages = held<kChain>(a.b_stages), p_plane_off = held<kChain>(a.a_plane_off), p_two = held<kChain>(a.in_planes == 2 && !a.a_split);
    const uint32_t p_stage = held<kChain>(a.stage_bytes), p_tx = held<kChain>(a.tx_bytes), p_stages = held<kChain>(a.stages), p_ks = held<kChain>(a.ksize), p_st = held<kChain>(a.stride);
    if (a.patch) {
      // K order = (channel block, tap): one patch load per channel block, nine weight tiles streamed through the B ring
      int sa = 0, pha = 0, sb = 0, phb = 0;
      for (int cbi = 0; cbi < num_kb; ++cbi) {
        const int cb = kb0 + cbi;
        mbar_wait(aempty_bar(sa), pha ^ 1);
        const uint32_t sta = smem_base + sa * p_a_stage;
        {
          mbar_expect_tx(afull_bar(sa), load ? p_a_tx : 0);
          if (load) {
            tma_load_5d(sta, &maps.in[0], afull_bar(sa), cb * p_ck, w0 - 1, h0 - 1, b, 0);
            if (p_two) tma_load_5d(sta + p_plane_off, &maps.in[0], afull_bar(sa), cb * p_ck, w0 - 1, h0 - 1, b, 1);
          }
          if (cbi < 96) stamp(8 + cbi);
        }
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(empty_bar(sb), phb ^ 1);
          {
            mbar_expect_tx(full_bar(sb), load ? p_b_tx : 0);
            if (load) tma_load_3d(smem_base + p_b_ring + sb * p_b_stage, &maps.w, full_bar(sb), (tap * p_kbt + cb) * p_ck, n0, 0);
          }
          if (++sb == static_cast<int>(p_b_stages)) { sb = 0; phb ^= 1; }
        }
        if (++sa == 2) { sa = 0; pha ^= 1; }
      }
    } else {
      int s = 0, ph = 0, tap = kb0 / a.kb_per_tap, cb = kb0 - tap * a.kb_per_tap;
      const uint32_t b_off = held<kChain>(a.a_region_bytes);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1);
        int map = 0, dh = 0, dw = 0;
        if (p_ks == 3) {
          const int kh = tap / 3, kw = tap - kh * 3;
          if (p_st == 1) { dh = kh - 1; dw = kw - 1; }
          else { map = ((kh == 1) ? 0 : 2) + ((kw == 1) ? 0 : 1); dh = (kh == 0) ? -1 : 0; dw = (kw == 0) ? -1 : 0; }
        } else if (p_ks == 0) {
          dh = static_cast<int>((a.tap_dh >> (4 * tap)) & 15ull) - 8;
          dw = static_cast<int>((a.tap_dw >> (4 * tap)) & 15ull) - 8;
        }
        const uint32_t st = smem_base + s * p_stage;
        {
          mbar_expect_tx(full_bar(s), load ? p_tx : 0);
          if (load) {
            tma_load_5d(st, &maps.in[map], full_bar(s), cb * p_ck, w0 + dw, h0 + dh, b, 0);
            if (p_two) tma_load_5d(st + p_plane_off, &maps.in[map], full_bar(s), cb * p_ck, w0 + dw, h0 + dh, b, 1);
            tma_load_3d(st + b_off, &maps.w, full_bar(s), (kb0 + kb) * p_ck, n0, 0);   // box covers both weight planes
          }
          if (kb < 96) stamp(8 + kb);
        }
        if (++cb == static_cast<int>(p_kbt)) { cb = 0; ++tap; }
        if (++s == static_cast<int>(p_stages)) { s = 0; ph ^= 1; }
      }
    }
  }
  // ===================== MMA issuers =====================
  // Measured on B200: one small tcgen05.mma (N <= 128, 32 bytes of K) occupies the tensor pipe for ~75-100 cycles whatever
  // its N, plus ~350 cycles per k-block for the barrier round trip of the issuing thread.  Hence (i) 3xTF32 stacks
  // W_hi and W_lo along N -- the two weight planes are adjacent in shared memory, so A_hi x [W_hi; W_lo] is ONE MMA of
  // N = 2*Nt yielding the main product and one cross term -- leaving A_lo x W_hi as the only other MMA per k-step, and
  // (ii) the two MMA streams are issued by two threads (warp 1 and lane 0 of the first epilogue warp, idle until the
  // accumulators are complete); bf16 deals the k-steps to the two issuers round-robin.  Each issuer rotates through its
  // own accumulators (dependent MMAs on one accumulator serialise, and fp32 accumulation in the tensor core truncates).
  // The issuing code is written warp-uniform (the whole warp walks the loop, one elected lane executes each tcgen05
  // instruction) with 32-bit descriptor arithmetic: tcgen05.mma / tcgen05.commit take uniform-register operands, and code
  // under a `lane == 0` branch makes the compiler move every operand through an elect / R2UR loop and 64-bit adds -- measured
  // ~450 cycles of issue overhead per MMA, several times the 64-128 cycles the MMA occupies the tensor pipe.
  if (warp >= 1 && warp - 1 < a.n_iss && elect_one()) {   // ONE lane runs the whole issue loop (see umma32_one)
    const int q = warp - 1;
    const uint32_t i_ck = held<kChain>(a.ck_bytes);
    const int ksteps = i_ck / 32;  // one UMMA consumes 32 bytes of K per row (8 tf32 / 16 bf16)
    const uint32_t b_plane = a.Nt * i_ck;
    const int cnt = held<kChain>(a.iss_cnt[q]);
    const uint32_t col0 = held<kChain>(tmem_base + a.iss_col[q]), cstride = held<kChain>(a.iss_stride[q]), idesc = held<kChain>(a.iss_idesc[q]);
    const int kstart = held<kChain>(a.kstart[q]), kinc = held<kChain>(a.kinc[q]);
    const uint32_t a_off0 = held<kChain>(a.job_a[q][0] * a.a_plane_off), b_off0 = held<kChain>(a.job_b[q][0] * b_plane);
    const uint32_t a_off1 = held<kChain>(a.job_a[q][1] * a.a_plane_off), b_off1 = held<kChain>(a.job_b[q][1] * b_plane);
    const bool two = held<kChain>(a.n_jobs[q] == 2) != 0;
    const uint32_t dhi = held<kChain>(smem_desc_hi(i_ck));
    const uint32_t i_a_stage = held<kChain>(a.a_stage_bytes), i_b_ring = held<kChain>(a.b_ring_off), i_b_stage = held<kChain>(a.b_stage_bytes);
    const int i_b_stages = held<kChain>(a.b_stages), i_stages = held<kChain>(a.stages);
    const uint32_t i_row_step = held<kChain>((a.Wp - 2) * i_ck), i_stage = held<kChain>(a.stage_bytes), i_a_region = held<kChain>(a.a_region_bytes);
    const bool live = a.ablate != 1;
    const bool fast4 = held<kChain>(live && ksteps == 4 && kinc == 1 && !two && a.ablate != 8) != 0;   // YP_CONV_ABLATE=8: generic loop (A/B runs)
    const bool fast2 = held<kChain>(live && ksteps == 4 && kinc == 2 && !two && a.ablate != 8) != 0;   // two issuers share a stream: every other k-step
    const bool fast4two = held<kChain>(live && ksteps == 4 && kinc == 1 && two && cnt == 1 && a.ablate != 8) != 0;   // both cross terms of the un-stacked wide plan
    uint32_t used = 0;                   // bit r set = accumulator r of this issuer already holds a partial sum
    int s = 0, ph = 0, nxt = 0;
    if (a.patch) {
      int sa = 0, pha = 0;
      for (int cbi = 0; cbi < num_kb; ++cbi) {
        mbar_wait(afull_bar(sa), pha);
        tc_fence_after();
        if (q == 0 && cbi < 96) stamp(104 + cbi);
        const uint32_t pa = smem_base + sa * i_a_stage;
        uint32_t shift = 0;               // byte offset of the tap's window inside the patch: (kh * Wp + kw) rows
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sb = smem_base + i_b_ring + s * i_b_stage;
          uint32_t al0 = smem_desc_lo(pa + a_off0 + shift) + 2 * kstart, bl0 = smem_desc_lo(sb + b_off0) + 2 * kstart;
          uint32_t al1 = smem_desc_lo(pa + a_off1 + shift) + 2 * kstart, bl1 = smem_desc_lo(sb + b_off1) + 2 * kstart;
          if (fast4) {
            // common case (128-byte rows, one MMA per k-step, every k-step by this issuer): four MMAs fully unrolled so that the
            // operand set-up of MMA k+1 overlaps the issue of MMA k
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma32_one<kTf32>(col0 + nxt * cstride, al0 + 2 * k, bl0 + 2 * k, dhi, idesc, (used >> nxt) & 1u);
              used |= 1u << nxt;
              nxt = (nxt + 1 == cnt) ? 0 : nxt + 1;
            }
          } else if (fast2) {
#pragma unroll
            for (int k = 0; k < 4; k += 2) {
              umma32_one<kTf32>(col0 + nxt * cstride, al0 + 2 * k, bl0 + 2 * k, dhi, idesc, (used >> nxt) & 1u);
              used |= 1u << nxt;
              nxt = (nxt + 1 == cnt) ? 0 : nxt + 1;
            }
          } else if (fast4two) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma32_one<kTf32>(col0, al0 + 2 * k, bl0 + 2 * k, dhi, idesc, used);
              umma32_one<kTf32>(col0, al1 + 2 * k, bl1 + 2 * k, dhi, idesc, 1u);
              used = 1u;
            }
          } else
          for (int k = kstart; k < ksteps && live; k += kinc) {
            umma32_one<kTf32>(col0 + nxt * cstride, al0, bl0, dhi, idesc, (used >> nxt) & 1u);
            if (two) umma32_one<kTf32>(col0 + nxt * cstride, al1, bl1, dhi, idesc, 1u);
            used |= 1u << nxt;
            nxt = (nxt + 1 == cnt) ? 0 : nxt + 1;
            al0 += 2 * kinc; bl0 += 2 * kinc; al1 += 2 * kinc; bl1 += 2 * kinc;
          }
          umma_commit(empty_bar(s));
          if (++s == i_b_stages) { s = 0; ph ^= 1; }
          shift += (tap % 3 == 2) ? i_row_step : i_ck;
        }
        umma_commit(aempty_bar(sa));
        if (q == 0 && cbi < 96) stamp(200 + cbi);
        if (++sa == 2) { sa = 0; pha ^= 1; }
      }
    } else {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (q == 0 && kb < 96) stamp(104 + kb);
        const uint32_t sa = smem_base + s * i_stage;
        const uint32_t sb = sa + i_a_region;
        uint32_t al0 = smem_desc_lo(sa + a_off0) + 2 * kstart, bl0 = smem_desc_lo(sb + b_off0) + 2 * kstart;
        uint32_t al1 = smem_desc_lo(sa + a_off1) + 2 * kstart, bl1 = smem_desc_lo(sb + b_off1) + 2 * kstart;
        if (fast4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma32_one<kTf32>(col0 + nxt * cstride, al0 + 2 * k, bl0 + 2 * k, dhi, idesc, (used >> nxt) & 1u);
            used |= 1u << nxt;
            nxt = (nxt + 1 == cnt) ? 0 : nxt + 1;
          }
        } else if (fast2) {
#pragma unroll
          for (int k = 0; k < 4; k += 2) {
            umma32_one<kTf32>(col0 + nxt * cstride, al0 + 2 * k, bl0 + 2 * k, dhi, idesc, (used >> nxt) & 1u);
            used |= 1u << nxt;
            nxt = (nxt + 1 == cnt) ? 0 : nxt + 1;
          }
        } else if (fast4two) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma32_one<kTf32>(col0, al0 + 2 * k, bl0 + 2 * k, dhi, idesc, used);
            umma32_one<kTf32>(col0, al1 + 2 * k, bl1 + 2 * k, dhi, idesc, 1u);
            used = 1u;
          }
        } else
        for (int k = kstart; k < ksteps && live; k += kinc) {
          umma32_one<kTf32>(col0 + nxt * cstride, al0, bl0, dhi, idesc, (used >> nxt) & 1u);
          if (two) umma32_one<kTf32>(col0 + nxt * cstride, al1, bl1, dhi, idesc, 1u);
          used |= 1u << nxt;
          nxt = (nxt + 1 == cnt) ? 0 : nxt + 1;
          al0 += 2 * kinc; bl0 += 2 * kinc; al1 += 2 * kinc; bl1 += 2 * kinc;
        }
        umma_commit(empty_bar(s));  // one arrival per issuer: the stage is free when all their MMAs have retired
        if (q == 0 && kb < 96) stamp(200 + kb);
        if (++s == i_stages) { s = 0; ph ^= 1; }
      }
    }
    umma_commit(accum_bar);
    if (kChain && q == 0) stamp(400);
  }
  __syncwarp();
  // The epilogue is the only part that depends on the output format / store-chunk width: a generic lambda, instantiated once by the
  // stand-alone kernel (its template arguments) and once per (format, units) combination by a chained item, which selects it at run
  // time -- the chain kernel then carries ONE copy of the set-up / producer / issuer code (instruction-cache footprint).
  auto epilogue = [&](auto fmt_tag, auto units_tag) {
    constexpr int OUT_FMT_ = decltype(fmt_tag)::value;
    constexpr int UNITS_ = decltype(units_tag)::value;
    using TO = typename OutT<OUT_FMT_>::type;
    constexpr int CH = 16 * UNITS_;                // elements per staging row
    constexpr int ROWB = CH * (int)sizeof(TO);     // bytes per staging row (128 / 64 / 32)
    constexpr int G = NG > UNITS_ ? NG / UNITS_ : 1; // staging chunks worked on at the same time (every warp group has a unit)
    constexpr int NP = UNITS_ / PU;                // passes per chunk
    const int n_chunks = a.Nt / CH;
    // ===================== epilogue =====================
    // All warps take part (the producer and issuer warps join when their loops are done): NG = NT / 128 warps per TMEM lane
    // quarter, which split the 16-column units of the tile between them (warp group hf takes the units u with u % NG == hf).
    // With one warp per scheduler the epilogue was bound by single-warp instruction latency (~4 cycles per instruction);
    // layers that cannot fill the GPU anyway (one CTA per SM) run 16 warps = four per quarter, which halves it again.
    // accumulator columns [col, col+16) of this thread's row: sum of all TMEM sources (main products first)
    auto tmem_acc16 = [&](int col, float* v) {
      float t[4][16];
      tmem_ld16(taddr_row + a.src_col[0] + col, v);
      int j = 1;
      for (; j + 2 < a.n_src; j += 3) {   // batches of three loads in flight
        tmem_ld16(taddr_row + a.src_col[j] + col, t[0]);
        tmem_ld16(taddr_row + a.src_col[j + 1] + col, t[1]);
        tmem_ld16(taddr_row + a.src_col[j + 2] + col, t[2]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = ((v[i] + t[0][i]) + t[1][i]) + t[2][i];
      }
      for (; j < a.n_src; ++j) {
        tmem_ld16(taddr_row + a.src_col[j] + col, t[3]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += t[3][i];
      }
      tmem_ld_wait();
    };
    const int S = a.split_k;
    const long long tile_id = static_cast<long long>(bx) * gy + by;
    const long long n_tiles_all = static_cast<long long>(gx) * gy;
    // split-K: partial sums of the S CTAs of a tile, read back in fixed order z = 0..S-1 (deterministic)
    auto load_acc16 = [&](int col, float* v) {
      if (S == 1) { tmem_acc16(col, v); return; }
      const float* p = a.ws_partial + (tile_id * 128 + row) * a.Nt + col;
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.0f;
      const long long zstride = n_tiles_all * 128 * a.Nt;
      for (int z0 = 0; z0 < S; z0 += 4) {   // four slices (16 x 16-byte loads) in flight per thread; summed in z order
        float4 f[4][4];
#pragma unroll
        for (int zz = 0; zz < 4; ++zz) {
          const float4* p4 = reinterpret_cast<const float4*>(p + (z0 + zz) * zstride);
#pragma unroll
          for (int j = 0; j < 4; ++j) f[zz][j] = (z0 + zz < S) ? __ldcg(p4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int zz = 0; zz < 4; ++zz)
#pragma unroll
          for (int j = 0; j < 4; ++j) { v[j * 4] += f[zz][j].x; v[j * 4 + 1] += f[zz][j].y; v[j * 4 + 2] += f[zz][j].z; v[j * 4 + 3] += f[zz][j].w; }
      }
    };
    auto finish = [&](float acc, int col, float res) -> float {
      float v = acc + bias_s[col];
      const float sv = silu_fast(v);
      v = a.act == YP_ACT_SILU ? sv : v;
      return v + res;
    };

    if (!kChain) asm volatile("griddepcontrol.wait;" ::: "memory");     // residual / split-K workspace may be the previous kernel's output
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    if (et0) stamp(2);

    bool run_epilogue = a.ablate != 7;   // 7 = main loop only (timing experiment)
    if (S > 1 && a.ablate != 7) {
      // publish this CTA's partial tile, then the last CTA to arrive (per tile) reduces all of them and finishes
      float* mine = a.ws_partial + ((bz * n_tiles_all + tile_id) * 128 + row) * a.Nt;
      for (int u = hf; u < a.Nt / 16; u += NG) {
        float v[16];
        tmem_acc16(u * 16, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) __stcg(reinterpret_cast<float4*>(mine + u * 16) + j, make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]));
      }
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
      if (et0) {
        const int old = atomicAdd(a.ws_counter + tile_id, 1);
        const int last = old == S - 1;
        if (last) a.ws_counter[tile_id] = 0;   // ready for the next launch that uses this workspace
        *split_flag = last;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
      run_epilogue = *split_flag != 0;
      __threadfence();
    }
    if (run_epilogue && a.rowmin) {
      // descriptor matching: key(i, j) = bits(sqrt(2 - 2 clip(<d1_i, d2_j>))) << 32 | j, integer MIN over this tile's columns
      const int n_rows = a.n_rows ? min(*a.n_rows, a.Wo) : a.Wo;
      const int n_cols = a.n_cols ? *a.n_cols : (int)(gy * a.Nt);
      unsigned long long best = ~0ull;
      // Column minima of the same tile (key(i, j) with the ROW index in the low word: the match in the other direction, which used to
      // be a second pass over the transposed problem): the tile's distances go to shared memory column-major (the pipeline stages
      // are free: every MMA has retired), pitch 129 floats so that the row-wise writes and the column-wise scans are conflict-free.
      float* dist_s = reinterpret_cast<float*>(smem_gen);
      const bool row_ok = valid && ow < n_rows;
      for (int u = hf; u < a.Nt / 16; u += NG) {
        float v[16];
        load_acc16(u * 16, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int j = n0 + u * 16 + e;
          const float d = fminf(fmaxf(v[e], -1.0f), 1.0f);
          const float dist = sqrtf(__fsub_rn(2.0f, __fmul_rn(2.0f, d)));
          const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(dist)) << 32) | static_cast<unsigned int>(j + a.col_off);
          if (j < n_cols) best = min(best, key);
          if (a.col_key) dist_s[(u * 16 + e) * 129 + row] = row_ok ? dist : __int_as_float(0x7f800000);
        }
      }
      if (row_ok && best != ~0ull) atomicMin(a.row_key + ow, best);
      if (a.col_key) {
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        // thread t scans half a column: column t % Nt... two threads per column when NT >= 2 Nt, each over 64 rows
        const int per_col = NT / a.Nt >= 2 ? 2 : 1;
        for (int idx = threadIdx.x; idx < a.Nt * per_col; idx += NT) {
          const int c = idx / per_col, part = idx - c * per_col;
          const int r0 = part * (128 / per_col), r1 = r0 + 128 / per_col;
          unsigned long long cbest = ~0ull;
          for (int r = r0; r < r1; ++r) {
            const float dist = dist_s[c * 129 + r];
            // row r of the tile is pixel (w0 + r) of the [1, 1, N1, D] view (Ht = 1, Wt = 128 for such views)
            const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(dist)) << 32) | static_cast<unsigned int>(w0 + r);
            cbest = min(cbest, key);
          }
          if (n0 + c < n_cols && (cbest >> 32) != 0x7f800000ull) atomicMin(a.col_key + n0 + c, cbest);
        }
      }
      run_epilogue = false;
    }
    if (run_epilogue) {
    float inv_norm = 1.0f;
    if (a.l2norm) {                 // desc / ||desc||_2 over all Nt channels of the pixel (single N tile)
      float ss = 0.0f;
      for (int u = 0; u < a.Nt / 16; ++u) {
        float v[16];
        load_acc16(u * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float f = finish(v[i], u * 16 + i, 0.0f); ss = fmaf(f, f, ss); }
      }
      inv_norm = 1.0f / sqrtf(ss);
    }

    // Chunks (one staging row = CH columns each) are processed G at a time so that every warp group has a 16-column unit; the
    // 2 G staging sets alternate between consecutive groups of chunks (the TMA stores of one group drain under the next).
    for (int cg = 0; cg < n_chunks; cg += G) {
      const int set0 = ((cg / G) & 1) * G;
      // the staging sets of this group were drained by the TMA stores committed two groups ago (thread et0 waited)
      asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
      if (et0 && cg == 0) stamp(320);
#pragma unroll 1
      for (int it = hf; it < G * NP; it += NG) {     // this warp group's units of the G chunks
      const int gi = it / NP, ps = it - gi * NP;
      const int c = cg + gi;
      if (c >= n_chunks) break;
      const uint32_t stg = smem_base + (set0 + gi) * a.staging_set_bytes;
      float v[PE];
      const int col0 = c * CH + ps * PE;
      if (!(res_pre && col0 == hf * 16)) load_res(OUT_FMT_, col0, res);
      if (et0 && c == 0 && ps == 0) stamp(321);
      if (a.ablate != 3) load_acc16(col0, v);
      if (et0 && c == 0 && ps == 0) stamp(322);
      if (a.act == YP_ACT_SILU) {
#pragma unroll
        for (int i = 0; i < PE; ++i) v[i] = (silu_fast(v[i] + bias_s[col0 + i]) + res[i]) * inv_norm;
      } else {
#pragma unroll
        for (int i = 0; i < PE; ++i) v[i] = (v[i] + bias_s[col0 + i] + res[i]) * inv_norm;
      }
      if (et0 && c < 8) stamp(300 + 4 * c);
      constexpr int VP = PE * (int)sizeof(TO) / 16;     // 16-byte vectors this pass contributes to the staging row
      const int j0 = ps * VP;
      if (!in_tile || a.ablate == 4) {
        // halo / padding row: nothing to stage
      } else if (OUT_FMT_ == YP_FMT_F32X2) {
#pragma unroll
        for (int j = 0; j < VP; ++j) {
          float hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) { hi[e] = tf32_round(v[j * 4 + e]); lo[e] = tf32_round(v[j * 4 + e] - hi[e]); }
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(swz_addr(stg, srow, j0 + j, ROWB)), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(swz_addr(stg + 128 * ROWB, srow, j0 + j, ROWB)), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]) : "memory");
        }
      } else if (OUT_FMT_ == YP_FMT_F32) {
#pragma unroll
        for (int j = 0; j < VP; ++j)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(swz_addr(stg, srow, j0 + j, ROWB)), "f"(v[j * 4]), "f"(v[j * 4 + 1]), "f"(v[j * 4 + 2]), "f"(v[j * 4 + 3]) : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < VP; ++j) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j * 8 + 2 * e], v[j * 8 + 2 * e + 1]);
            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(swz_addr(stg, srow, j0 + j, ROWB)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
      }
      }  // units
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
      if (et0 && cg < 8) stamp(301 + 4 * cg);
      if (et0 && a.ablate != 5) {
        for (int gi = 0; gi < G && cg + gi < n_chunks; ++gi)
          for (int m = 0; m < a.n_out_maps; ++m)
            for (int pl = 0; pl < a.out_planes; ++pl)
              tma_store_5d(&maps.out[m], smem_base + (set0 + gi) * a.staging_set_bytes + pl * 128 * ROWB, n0 + (cg + gi) * CH, w0, h0, b, pl);
        tma_store_commit();
        tma_store_wait_read<1>();  // everything but the group just committed has finished reading smem
        if (cg < 8) stamp(302 + 4 * cg);
      }
    }
    if (et0) {
      if (kChain) tma_store_wait_all();   // the stores have been written, not only read out of the staging buffers
      else tma_store_wait_read<0>();
      stamp(3);
    }
    }  // run_epilogue
  };
  if constexpr (kChain) {
    switch (chain_variant) {
      case 0: epilogue(std::integral_constant<int, YP_FMT_F32X2>{}, std::integral_constant<int, 2>{}); break;
      case 1: epilogue(std::integral_constant<int, YP_FMT_F32X2>{}, std::integral_constant<int, 1>{}); break;
      case 2: epilogue(std::integral_constant<int, YP_FMT_F32>{}, std::integral_constant<int, 2>{}); break;
      default: epilogue(std::integral_constant<int, YP_FMT_F32>{}, std::integral_constant<int, 1>{}); break;
    }
  } else {
    epilogue(std::integral_constant<int, OUT_FMT>{}, std::integral_constant<int, UNITS>{});
  }

  }  // ablate != 6
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(4);
  if (!kChain && warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
  if (threadIdx.x == 32) stamp(5);
  if (!kChain && a.dbg && threadIdx.x == 0 && cta_lin < 20000) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg[514 + 3 * cta_lin] = static_cast<long long>(t);
  }
}

template <int OUT_FMT, int UNITS, bool kTf32, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1) conv_tc_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  conv_tile<OUT_FMT, UNITS, kTf32, NT, false>(maps, a, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.x, gridDim.y, smem_raw, 0u, true, [] {});
}

// ---------------------------------------------------------------------------------------------
// Persistent variant for layers with several waves of tiles (bf16; no split-K / L2-norm / row-min): one CTA per SM walks the
// tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the accumulators are double-buffered in TMEM (2 x 256 columns), so the
// epilogue of tile i (warps 4-11: TMEM -> registers -> bias / SiLU / residual -> swizzled staging -> TMA store) runs under the
// K loop of tile i+1 (warp 0 = TMA producer, warps 1-2 = MMA issuers, all continuing through the same mbarrier rings), and the
// prologue (barrier init, TMEM allocation, descriptor prefetch) is paid once per SM instead of once per tile.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int OUT_FMT, int UNITS, bool kTf32>
__global__ void __launch_bounds__(kPersistThreads, 1) conv_tc_persist_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int per_img = a.tiles_w * a.tiles_h;
  const int num_kb = a.patch ? a.kb_per_tap : a.n_taps * a.kb_per_tap;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const uint32_t bar_base = smem_base + a.bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto afull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
  auto accf_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 4 + s); };
  auto acce_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 6 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 8);

  if (warp == 0 && lane == 0) {
    const int n_in = a.stride == 2 ? 4 : 1;
    for (int i = 0; i < n_in; ++i) tma_prefetch_desc(&maps.in[i]);
    tma_prefetch_desc(&maps.w);
    for (int i = 0; i < a.n_out_maps; ++i) tma_prefetch_desc(&maps.out[i]);
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), a.n_iss); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(afull_bar(s), 1); mbar_init(aempty_bar(s), a.n_iss);
      mbar_init(accf_bar(s), a.n_iss); mbar_init(acce_bar(s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  // tile t -> (image b, patch origin h0 / w0, first output channel n0); the N tiles of one patch are consecutive
  auto decode = [&](int t, int& b, int& h0, int& w0, int& n0) {
    const int nt = t % a.n_tiles_n, m = t / a.n_tiles_n;
    b = m / per_img;
    const int trem = m - b * per_img;
    const int th = trem / a.tiles_w, tw = trem - th * a.tiles_w;
    h0 = th * a.Ht; w0 = tw * a.Wt; n0 = nt * a.Nt;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int s = 0, ph = 0, sa = 0, pha = 0;
    for (int t = blockIdx.x; t < a.tiles_total; t += gridDim.x) {
      int b, h0, w0, n0;
      decode(t, b, h0, w0, n0);
      if (a.patch) {
        for (int cb = 0; cb < num_kb; ++cb) {
          mbar_wait(aempty_bar(sa), pha ^ 1);
          const uint32_t sta = smem_base + sa * a.a_stage_bytes;
          if (elect_one()) {
            mbar_expect_tx(afull_bar(sa), a.a_tx);
            tma_load_5d(sta, &maps.in[0], afull_bar(sa), cb * a.ck_elems, w0 - 1, h0 - 1, b, 0);
            if (a.in_planes == 2 && !a.a_split) tma_load_5d(sta + a.a_plane_off, &maps.in[0], afull_bar(sa), cb * a.ck_elems, w0 - 1, h0 - 1, b, 1);
          }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(empty_bar(s), ph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(full_bar(s), a.b_tx);
              tma_load_3d(smem_base + a.b_ring_off + s * a.b_stage_bytes, &maps.w, full_bar(s), (tap * a.kb_per_tap + cb) * a.ck_elems, n0, 0);
            }
            if (++s == a.b_stages) { s = 0; ph ^= 1; }
          }
          if (++sa == 2) { sa = 0; pha ^= 1; }
        }
      } else {
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1);
          int map = 0, dh = 0, dw = 0;
          if (a.ksize == 3) {
            const int kh = tap / 3, kw = tap - kh * 3;
            if (a.stride == 1) { dh = kh - 1; dw = kw - 1; }
            else { map = ((kh == 1) ? 0 : 2) + ((kw == 1) ? 0 : 1); dh = (kh == 0) ? -1 : 0; dw = (kw == 0) ? -1 : 0; }
          } else if (a.ksize == 0) {
            dh = static_cast<int>((a.tap_dh >> (4 * tap)) & 15ull) - 8;
            dw = static_cast<int>((a.tap_dw >> (4 * tap)) & 15ull) - 8;
          }
          const uint32_t st = smem_base + s * a.stage_bytes;
          if (elect_one()) {
            mbar_expect_tx(full_bar(s), a.tx_bytes);
            tma_load_5d(st, &maps.in[map], full_bar(s), cb * a.ck_elems, w0 + dw, h0 + dh, b, 0);
            if (a.in_planes == 2 && !a.a_split) tma_load_5d(st + a.a_plane_off, &maps.in[map], full_bar(s), cb * a.ck_elems, w0 + dw, h0 + dh, b, 1);
            tma_load_3d(st + a.a_region_bytes, &maps.w, full_bar(s), kb * a.ck_elems, n0, 0);
          }
          if (++cb == a.kb_per_tap) { cb = 0; ++tap; }
          if (++s == a.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp - 1 < a.n_iss) {
    // ===================== MMA issuers =====================
    const int q = warp - 1;
    const int ksteps = a.ck_bytes / 32;
    const uint32_t b_plane = a.Nt * a.ck_bytes;
    const int cnt = a.iss_cnt[q];
    const uint32_t cstride = a.iss_stride[q], idesc = a.iss_idesc[q];
    const int kstart = a.kstart[q], kinc = a.kinc[q];
    const uint32_t a_off0 = a.job_a[q][0] * a.a_plane_off, b_off0 = a.job_b[q][0] * b_plane;
    const uint32_t a_off1 = a.job_a[q][1] * a.a_plane_off, b_off1 = a.job_b[q][1] * b_plane;
    const bool two = a.n_jobs[q] == 2;
    const uint32_t dhi = smem_desc_hi(a.ck_bytes);
    int s = 0, ph = 0, sa = 0, pha = 0, it = 0;
    for (int t = blockIdx.x; t < a.tiles_total; t += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(acce_bar(buf), ((it >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t col0 = tmem_base + buf * a.buf_cols + a.iss_col[q];
      uint32_t used = 0;
      int nxt = 0;
      if (a.patch) {
        for (int cb = 0; cb < num_kb; ++cb) {
          mbar_wait(afull_bar(sa), pha);
          tc_fence_after();
          const uint32_t pa = smem_base + sa * a.a_stage_bytes;
          uint32_t shift = 0;
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t sb = smem_base + a.b_ring_off + s * a.b_stage_bytes;
            uint32_t al0 = smem_desc_lo(pa + a_off0 + shift) + 2 * kstart, bl0 = smem_desc_lo(sb + b_off0) + 2 * kstart;
            uint32_t al1 = smem_desc_lo(pa + a_off1 + shift) + 2 * kstart, bl1 = smem_desc_lo(sb + b_off1) + 2 * kstart;
            for (int k = kstart; k < ksteps; k += kinc) {
              umma32<kTf32>(col0 + nxt * cstride, al0, bl0, dhi, idesc, (used >> nxt) & 1u);
              if (two) umma32<kTf32>(col0 + nxt * cstride, al1, bl1, dhi, idesc, 1u);
              used |= 1u << nxt;
              nxt = (nxt + 1 == cnt) ? 0 : nxt + 1;
              al0 += 2 * kinc; bl0 += 2 * kinc; al1 += 2 * kinc; bl1 += 2 * kinc;
            }
            umma_commit_elect(empty_bar(s));
            if (++s == a.b_stages) { s = 0; ph ^= 1; }
            shift += (tap % 3 == 2) ? (a.Wp - 2) * a.ck_bytes : a.ck_bytes;
          }
          umma_commit_elect(aempty_bar(sa));
          if (++sa == 2) { sa = 0; pha ^= 1; }
        }
      } else {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sA = smem_base + s * a.stage_bytes;
          const uint32_t sB = sA + a.a_region_bytes;
          uint32_t al0 = smem_desc_lo(sA + a_off0) + 2 * kstart, bl0 = smem_desc_lo(sB + b_off0) + 2 * kstart;
          uint32_t al1 = smem_desc_lo(sA + a_off1) + 2 * kstart, bl1 = smem_desc_lo(sB + b_off1) + 2 * kstart;
          for (int k = kstart; k < ksteps; k += kinc) {
            umma32<kTf32>(col0 + nxt * cstride, al0, bl0, dhi, idesc, (used >> nxt) & 1u);
            if (two) umma32<kTf32>(col0 + nxt * cstride, al1, bl1, dhi, idesc, 1u);
            used |= 1u << nxt;
            nxt = (nxt + 1 == cnt) ? 0 : nxt + 1;
            al0 += 2 * kinc; bl0 += 2 * kinc; al1 += 2 * kinc; bl1 += 2 * kinc;
          }
          umma_commit_elect(empty_bar(s));
          if (++s == a.stages) { s = 0; ph ^= 1; }
        }
      }
      umma_commit_elect(accf_bar(buf));                   // this issuer's MMAs of the tile have retired
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps: two per TMEM lane quarter, splitting the 16-column units) =====================
    using TO = typename OutT<OUT_FMT>::type;
    constexpr int CH = 16 * UNITS;
    constexpr int ROWB = CH * (int)sizeof(TO);
    constexpr int VP = 16 * (int)sizeof(TO) / 16;          // 16-byte vectors per 16-column unit
    const int q = warp & 3;
    const int hf = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int rdiv = a.patch ? a.Wp : a.Wt;
    const int rh = row / rdiv, rw = row - rh * rdiv;
    const bool in_tile = rh < a.Ht && rw < a.Wt;
    const int srow = in_tile ? rh * a.Wt + rw : 127;
    const bool et0 = (threadIdx.x == 128);
    const int n_chunks = a.Nt / CH;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int it = 0, cg = 0;                                     // cg: running chunk counter (selects the staging set)
    for (int t = blockIdx.x; t < a.tiles_total; t += gridDim.x, ++it) {
      int b, h0, w0, n0;
      decode(t, b, h0, w0, n0);
      const int buf = it & 1;
      const int oh = h0 + rh, ow = w0 + rw;
      const bool valid = in_tile && oh < a.Ho && ow < a.Wo;
      const bool has_res = a.res_base != nullptr && valid;
      const long long res_off = ((static_cast<long long>(b) * a.Ho + oh) * a.Wo + ow) * a.res_pix + n0;
      const uint32_t taddr_row = tmem_base + buf * a.buf_cols + (static_cast<uint32_t>(q * 32) << 16);
      mbar_wait(accf_bar(buf), (it >> 1) & 1);
      tc_fence_after();
      for (int c = 0; c < n_chunks; ++c, ++cg) {
        const uint32_t stg = smem_base + a.stg_off + (cg & 1) * a.staging_set_bytes;
        asm volatile("bar.sync 2, 256;" ::: "memory");      // staging set (cg & 1) was drained (et0 waited for the store of chunk cg-2)
#pragma unroll 1
        for (int ps = 0; ps < UNITS; ++ps) {
          if (((c * UNITS + ps) & 1) != hf) continue;     // the other warp of this lane quarter takes this unit
          const int col0 = c * CH + ps * 16;
          float v[16], r[16];
          tmem_ld16(taddr_row + a.src_col[0] + col0, v);
          if (a.n_src > 1) {
            float t2[16];
            tmem_ld16(taddr_row + a.src_col[1] + col0, t2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += t2[i];
          } else {
            tmem_ld_wait();
          }
          if (has_res) {
            if (OUT_FMT == YP_FMT_BF16) {
              const uint4* p = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(a.res_base) + res_off + col0);
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint4 u = __ldg(p + j);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); r[j * 8 + 2 * e] = f.x; r[j * 8 + 2 * e + 1] = f.y; }
              }
            } else {
              const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(a.res_base) + res_off + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float4 f = __ldg(p + j); r[j * 4] = f.x; r[j * 4 + 1] = f.y; r[j * 4 + 2] = f.z; r[j * 4 + 3] = f.w; }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = 0.0f;
          }
          if (a.bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __ldg(a.bias + n0 + col0 + i);
          }
          if (a.act == YP_ACT_SILU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = silu_fast(v[i]) + r[i];
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += r[i];
          }
          if (in_tile) {
            if (OUT_FMT == YP_FMT_BF16) {
#pragma unroll
              for (int j = 0; j < VP; ++j) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j * 8 + 2 * e], v[j * 8 + 2 * e + 1]);
                  pk[e] = *reinterpret_cast<uint32_t*>(&h2);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(swz_addr(stg, srow, ps * VP + j, ROWB)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
              }
            } else {
#pragma unroll
              for (int j = 0; j < VP; ++j)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(swz_addr(stg, srow, ps * VP + j, ROWB)), "f"(v[j * 4]), "f"(v[j * 4 + 1]), "f"(v[j * 4 + 2]), "f"(v[j * 4 + 3]) : "memory");
            }
          }
        }
        if (c == n_chunks - 1) tc_fence_before();           // last TMEM read of this tile is done (tcgen05.wait::ld above)
        fence_proxy_async_smem();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (et0) {
          if (c == n_chunks - 1) mbar_arrive(acce_bar(buf)); // hand the accumulator buffer back to the issuers
          for (int m = 0; m < a.n_out_maps; ++m) tma_store_5d(&maps.out[m], stg, n0 + c * CH, w0, h0, b, 0);
          tma_store_commit();
          tma_store_wait_read<1>();
        }
      }
    }
    if (et0) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// 3xTF32 variant with DRAINED accumulators (a.persist == 2; the throughput plan, YP_TILE_WIDE).
//
// The tensor core adds into its fp32 accumulator with truncation, so the error of a tile grows with the number of MMAs chained on one
// accumulator; conv_tile() bounds the chains by rotating over many TMEM accumulators, which caps the N tile (TMEM columns) and forces
// split-K on deep layers.  Here the chains are bounded at ANY tile width: the main product A_hi x W_hi accumulates for one ROUND
// (a.drain_units pipeline units = ~16 MMAs) into one of two TMEM buffers; the epilogue warps -- idle during the main loop anyway --
// then pull the buffer into registers (tcgen05.ld) and add it to their running fp32 sums with round-to-nearest, while the issuer fills
// the other buffer.  The two cross terms (A_lo x W_hi, A_hi x W_lo; ~2^-11 of the result, their truncation error is negligible)
// accumulate over the whole K in a third accumulator that is read once per tile.  TMEM: main 2 x Nt + cross 2 x Nt columns
// (the cross accumulator alternates between two buffers per TILE), Nt <= 128; for Nt <= 64 the round buffers are 2 Nt wide and also
// take the A_hi x W_lo cross term (one stacked MMA of N = 2 Nt, see the issuer).
//
// Like conv_tc_persist_kernel the CTA is persistent (one per SM, tiles t = blockIdx.x, +gridDim.x, ...): warp 0 = TMA producer,
// warp 1 = main issuer, warp 2 = cross issuer, warps 4-11 = drain + epilogue (two per TMEM lane quarter; the warp pair splits the
// 16-column units of the tile by parity, so a thread keeps Nt / 2 running sums in registers).  The epilogue of tile i (bias / SiLU /
// residual / (hi, lo) split / staging / TMA stores) overlaps the first rounds of tile i+1.
// ---------------------------------------------------------------------------------------------
// (144 registers per thread instead of the 168 the compiler would take: 384 x 144 leaves a sixth of the register file to the small
// post-processing kernels of other frames in flight; measured 2461 -> 2482 frames/s, 128 registers: 2414)
template <int OUT_FMT>
__global__ void __maxnreg__(144) conv_tc_drain_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int per_img = a.tiles_w * a.tiles_h;
  const int num_kb = a.patch ? a.kb_per_tap : a.n_taps * a.kb_per_tap;
  const int units_per_tile = a.patch ? 9 * num_kb : num_kb;      // pipeline units (one weight stage each) of a tile
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const uint32_t bar_base = smem_base + a.bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto afull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
  auto accf_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 4 + s); };
  auto acce_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 6 + s); };
  auto crossf_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 8 + s); };
  auto crosse_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 10 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 12);

  if (warp == 0 && lane == 0) {
    const int n_in = a.stride == 2 ? 4 : 1;
    for (int i = 0; i < n_in; ++i) tma_prefetch_desc(&maps.in[i]);
    tma_prefetch_desc(&maps.w);
    for (int i = 0; i < a.n_out_maps; ++i) tma_prefetch_desc(&maps.out[i]);
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 2); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(afull_bar(s), 1); mbar_init(aempty_bar(s), 2);
      mbar_init(accf_bar(s), 1); mbar_init(acce_bar(s), 8);
      mbar_init(crossf_bar(s), 1); mbar_init(crosse_bar(s), 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  auto decode = [&](int t, int& b, int& h0, int& w0, int& n0) {
    const int nt = t % a.n_tiles_n, m = t / a.n_tiles_n;
    b = m / per_img;
    const int trem = m - b * per_img;
    const int th = trem / a.tiles_w, tw = trem - th * a.tiles_w;
    h0 = th * a.Ht; w0 = tw * a.Wt; n0 = nt * a.Nt;
  };

  if (warp == 0) {
    // ===================== TMA producer (as in conv_tc_persist_kernel) =====================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int s = 0, ph = 0, sa = 0, pha = 0;
    for (int t = blockIdx.x; t < a.tiles_total; t += gridDim.x) {
      int b, h0, w0, n0;
      decode(t, b, h0, w0, n0);
      if (a.patch) {
        for (int cb = 0; cb < num_kb; ++cb) {
          mbar_wait(aempty_bar(sa), pha ^ 1);
          const uint32_t sta = smem_base + sa * a.a_stage_bytes;
          if (elect_one()) {
            mbar_expect_tx(afull_bar(sa), a.a_tx);
            tma_load_5d(sta, &maps.in[0], afull_bar(sa), cb * a.ck_elems, w0 - 1, h0 - 1, b, 0);
            tma_load_5d(sta + a.a_plane_off, &maps.in[0], afull_bar(sa), cb * a.ck_elems, w0 - 1, h0 - 1, b, 1);
          }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(empty_bar(s), ph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(full_bar(s), a.b_tx);
              tma_load_3d(smem_base + a.b_ring_off + s * a.b_stage_bytes, &maps.w, full_bar(s), (tap * a.kb_per_tap + cb) * a.ck_elems, n0, 0);
            }
            if (++s == a.b_stages) { s = 0; ph ^= 1; }
          }
          if (++sa == 2) { sa = 0; pha ^= 1; }
        }
      } else {
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1);
          int map = 0, dh = 0, dw = 0;
          if (a.ksize == 3) {
            const int kh = tap / 3, kw = tap - kh * 3;
            if (a.stride == 1) { dh = kh - 1; dw = kw - 1; }
            else { map = ((kh == 1) ? 0 : 2) + ((kw == 1) ? 0 : 1); dh = (kh == 0) ? -1 : 0; dw = (kw == 0) ? -1 : 0; }
          }
          const uint32_t st = smem_base + s * a.stage_bytes;
          if (elect_one()) {
            mbar_expect_tx(full_bar(s), a.tx_bytes);
            tma_load_5d(st, &maps.in[map], full_bar(s), cb * a.ck_elems, w0 + dw, h0 + dh, b, 0);
            if (!a.a_split) tma_load_5d(st + a.a_plane_off, &maps.in[map], full_bar(s), cb * a.ck_elems, w0 + dw, h0 + dh, b, 1);
            tma_load_3d(st + a.a_region_bytes, &maps.w, full_bar(s), kb * a.ck_elems, n0, 0);
          }
          if (++cb == a.kb_per_tap) { cb = 0; ++tap; }
          if (++s == a.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if ((warp == 1 || warp == 2) && elect_one()) {
    // ===================== MMA issuers: warp 1 = main product (rounds, two TMEM buffers), warp 2 = both cross terms =====================
    const bool is_main = warp == 1;
    const int ksteps = a.ck_bytes / 32;
    const uint32_t b_plane = a.Nt * a.ck_bytes;
    // a.buf_cols = columns of one round buffer: Nt (main product only; both cross terms go to the cross accumulator) or, for
    // Nt <= 64, 2 Nt: the two weight planes are adjacent in shared memory, so A_hi x [W_hi; W_lo] is ONE MMA of N = 2 Nt that yields
    // the main product and one cross term side by side (both are drained into the same running sums), and the cross issuer is left
    // with A_lo x W_hi -- two MMAs per k-step instead of three (small MMAs cost ~64 cycles whatever their N).
    const bool stacked = a.buf_cols == 2 * a.Nt;
    const uint32_t idesc = a.iss_idesc[is_main ? 0 : 1];
    const uint32_t dhi = smem_desc_hi(a.ck_bytes);
    const uint32_t main_col = tmem_base, cross_col = tmem_base + 2 * a.buf_cols;
    int s = 0, ph = 0, sa = 0, pha = 0, tile_it = 0;
    int rc = 0;          // rounds issued so far (main issuer); buffer = rc & 1
    for (int t = blockIdx.x; t < a.tiles_total; t += gridDim.x, ++tile_it) {
      const int tb = tile_it & 1;
      uint32_t acc = 0;                       // accumulate flag of the next MMA of this issuer's current accumulator
      int u_in_round = 0;
      if (!is_main) {
        if (tile_it >= 2) { mbar_wait(crosse_bar(tb), ((tile_it >> 1) - 1) & 1); tc_fence_after(); }
      }
      auto begin_round = [&]() {
        const int buf = rc & 1;
        if (rc >= 2) { mbar_wait(acce_bar(buf), ((rc >> 1) - 1) & 1); tc_fence_after(); }
        acc = 0;
      };
      auto issue_unit = [&](uint32_t sa_addr, uint32_t sb_addr, bool last_unit) {
        if (is_main) {
          if (u_in_round == 0) begin_round();
          const uint32_t col = main_col + (rc & 1) * a.buf_cols;
          const uint32_t al = smem_desc_lo(sa_addr), bl = smem_desc_lo(sb_addr);
          for (int k = 0; k < ksteps; ++k) { umma32_one<true>(col, al + 2 * k, bl + 2 * k, dhi, idesc, acc); acc = 1u; }
          if (++u_in_round == a.drain_units || last_unit) { umma_commit(accf_bar(rc & 1)); ++rc; u_in_round = 0; }
        } else {
          const uint32_t col = cross_col + tb * a.Nt;
          const uint32_t al_hi = smem_desc_lo(sa_addr), al_lo = smem_desc_lo(sa_addr + a.a_plane_off);
          const uint32_t bl_hi = smem_desc_lo(sb_addr), bl_lo = smem_desc_lo(sb_addr + b_plane);
          for (int k = 0; k < ksteps; ++k) {
            umma32_one<true>(col, al_lo + 2 * k, bl_hi + 2 * k, dhi, idesc, acc);
            if (!stacked) umma32_one<true>(col, al_hi + 2 * k, bl_lo + 2 * k, dhi, idesc, 1u);
            acc = 1u;
          }
          if (last_unit) umma_commit(crossf_bar(tb));
        }
      };
      int unit = 0;
      if (a.patch) {
        for (int cb = 0; cb < num_kb; ++cb) {
          mbar_wait(afull_bar(sa), pha);
          tc_fence_after();
          const uint32_t pa = smem_base + sa * a.a_stage_bytes;
          uint32_t shift = 0;
          for (int tap = 0; tap < 9; ++tap, ++unit) {
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            issue_unit(pa + shift, smem_base + a.b_ring_off + s * a.b_stage_bytes, unit + 1 == units_per_tile);
            umma_commit(empty_bar(s));
            if (++s == a.b_stages) { s = 0; ph ^= 1; }
            shift += (tap % 3 == 2) ? (a.Wp - 2) * a.ck_bytes : a.ck_bytes;
          }
          umma_commit(aempty_bar(sa));
          if (++sa == 2) { sa = 0; pha ^= 1; }
        }
      } else {
        for (int kb = 0; kb < num_kb; ++kb, ++unit) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sA = smem_base + s * a.stage_bytes;
          issue_unit(sA, sA + a.a_region_bytes, unit + 1 == units_per_tile);
          umma_commit(empty_bar(s));
          if (++s == a.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== drain + epilogue (8 warps: two per TMEM lane quarter, units split by parity) =====================
    using TO = typename OutT<OUT_FMT>::type;
    constexpr int CH = 32;                                   // UNITS = 2: one staging row = 32 fp32 columns = 128 bytes
    constexpr int ROWB = CH * (int)sizeof(TO);
    const int q = warp & 3;
    const int hf = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int rdiv = a.patch ? a.Wp : a.Wt;
    const int rh = row / rdiv, rw = row - rh * rdiv;
    const bool in_tile = rh < a.Ht && rw < a.Wt;
    const int srow = in_tile ? rh * a.Wt + rw : 127;
    const bool et0 = (threadIdx.x == 128);
    const int n_chunks = a.Nt / CH;                          // <= 4; this thread owns unit 2 c + hf of chunk c
    const int rounds_per_tile = (units_per_tile + a.drain_units - 1) / a.drain_units;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int tile_it = 0, cg = 0, rc = 0;
    for (int t = blockIdx.x; t < a.tiles_total; t += gridDim.x, ++tile_it) {
      int b, h0, w0, n0;
      decode(t, b, h0, w0, n0);
      const int tb = tile_it & 1;
      float sums[4][16];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 16; ++i) sums[c][i] = 0.0f;
      // ---- drain the rounds of this tile
      for (int r = 0; r < rounds_per_tile; ++r, ++rc) {
        const int buf = rc & 1;
        mbar_wait(accf_bar(buf), (rc >> 1) & 1);
        tc_fence_after();
        const uint32_t base = tmem_base + lane_off + buf * a.buf_cols + hf * 16;
        const int halves = a.buf_cols == 2 * a.Nt ? 2 : 1;      // stacked round buffer: [main | A_hi x W_lo], both added to the same sums
        for (int hv = 0; hv < halves; ++hv) {
#pragma unroll
          for (int c0 = 0; c0 < 4; c0 += 2) {      // two units in flight (register budget: 64 running sums + 32 in flight)
            if (c0 < n_chunks) {
              float tv[2][16];
              tmem_ld16(base + hv * a.Nt + c0 * CH, tv[0]);
              if (c0 + 1 < n_chunks) tmem_ld16(base + hv * a.Nt + (c0 + 1) * CH, tv[1]);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) sums[c0][i] += tv[0][i];
              if (c0 + 1 < n_chunks) {
#pragma unroll
                for (int i = 0; i < 16; ++i) sums[c0 + 1][i] += tv[1][i];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acce_bar(buf));
      }
      // ---- cross terms of the tile
      mbar_wait(crossf_bar(tb), (tile_it >> 1) & 1);
      tc_fence_after();
      {
        const uint32_t base = tmem_base + lane_off + 2 * a.buf_cols + tb * a.Nt + hf * 16;
#pragma unroll
        for (int c0 = 0; c0 < 4; c0 += 2) {
          if (c0 < n_chunks) {
            float tv[2][16];
            tmem_ld16(base + c0 * CH, tv[0]);
            if (c0 + 1 < n_chunks) tmem_ld16(base + (c0 + 1) * CH, tv[1]);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) sums[c0][i] += tv[0][i];
            if (c0 + 1 < n_chunks) {
#pragma unroll
              for (int i = 0; i < 16; ++i) sums[c0 + 1][i] += tv[1][i];
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(crosse_bar(tb));
      }
      // ---- epilogue: bias, SiLU, residual, operand split, staging, TMA stores
      const int oh = h0 + rh, ow = w0 + rw;
      const bool valid = in_tile && oh < a.Ho && ow < a.Wo;
      const bool has_res = a.res_base != nullptr && valid;
      const long long res_off = ((static_cast<long long>(b) * a.Ho + oh) * a.Wo + ow) * a.res_pix + n0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < n_chunks) {
          const uint32_t stg = smem_base + a.stg_off + (cg & 1) * a.staging_set_bytes;
          asm volatile("bar.sync 2, 256;" ::: "memory");      // staging set (cg & 1) was drained (et0 waited for the stores of chunk cg-2)
          const int col0 = c * CH + hf * 16;
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = sums[c][i] + (a.bias ? __ldg(a.bias + n0 + col0 + i) : 0.0f);
          if (a.act == YP_ACT_SILU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = silu_fast(v[i]);
          }
          if (has_res) {
            const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(a.res_base) + res_off + col0);
            const float4* pl = reinterpret_cast<const float4*>(static_cast<const float*>(a.res_base) + res_off + a.res_plane + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 hi = __ldg(p + j);
              float4 lo = make_float4(0.f, 0.f, 0.f, 0.f);
              if (OUT_FMT == YP_FMT_F32X2) lo = __ldg(pl + j);
              v[j * 4] += hi.x + lo.x; v[j * 4 + 1] += hi.y + lo.y; v[j * 4 + 2] += hi.z + lo.z; v[j * 4 + 3] += hi.w + lo.w;
            }
          }
          if (in_tile) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (OUT_FMT == YP_FMT_F32X2) {
                float hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) { hi[e] = tf32_round(v[j * 4 + e]); lo[e] = tf32_round(v[j * 4 + e] - hi[e]); }
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(swz_addr(stg, srow, hf * 4 + j, ROWB)), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(swz_addr(stg + 128 * ROWB, srow, hf * 4 + j, ROWB)), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]) : "memory");
              } else {
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(swz_addr(stg, srow, hf * 4 + j, ROWB)), "f"(v[j * 4]), "f"(v[j * 4 + 1]), "f"(v[j * 4 + 2]), "f"(v[j * 4 + 3]) : "memory");
              }
            }
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (et0) {
            for (int m = 0; m < a.n_out_maps; ++m)
              for (int pl = 0; pl < a.out_planes; ++pl)
                tma_store_5d(&maps.out[m], stg + pl * 128 * ROWB, n0 + c * CH, w0, h0, b, pl);
            tma_store_commit();
            tma_store_wait_read<1>();
          }
          ++cg;
        }
      }
    }
    if (et0) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// Layer chain: one persistent kernel for a segment of the network (yp_conv_chain_*, see the header).  Batch-1 inference is a chain
// of ~60 small dependent layers; launched one by one each pays launch latency, TMEM allocation, descriptor prefetch and the
// grid-completion gap, and the GPU idles while the slowest CTA of a layer finishes.  Here one CTA per SM walks the operations in
// order; item i of operation l runs on CTA (cta_off_l + i) mod #CTAs exactly as conv_tile() would run it as a CTA of its own launch.
// Ordering between operations: done[l] counts the completed items of operation l (red.release.gpu after the item's stores have
// completed); an item first waits (ld.acquire.gpu, one thread, under the barrier set-up of the others) until every operation in its
// dependency list has reached its item count, then fences the async proxy (the data arrive through TMA).  Deadlock freedom: every
// CTA processes its items in list order, the list is a valid sequential order and all CTAs are co-resident (cooperative launch,
// grid = #SMs, one CTA per SM).  The last CTA to leave the kernel resets the counters for the next launch.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxChain = 40;      // operations per kernel (bounded by the 32 KB kernel-parameter space)
constexpr int kMaxChainDeps = 6;

struct ChainMeta {
  int type;                        // 0 = convolution, 1 = SPPF pooling
  int variant;                     // convolution: epilogue instantiation; pooling: index into ChainParams::pool
  int gx, gy, n_items, cta_off;
  int n_deps;
  int dep[kMaxChainDeps], dep_target[kMaxChainDeps];
};
struct ChainPool { YpView cat4; int C; int groups; };
struct ChainParams {
  int n_ops;
  long long* dbg;                  // optional per-item timeline (YP_CHAIN_DEBUG): [n_ops][#CTAs][16] globaltimer stamps
  ChainMeta meta[kMaxChain];
  ConvArgs args[kMaxChain];
  ChainPool pool[2];
};
static_assert(sizeof(ChainParams) <= 32000, "ChainParams exceeds the kernel parameter space");

__device__ __forceinline__ void chain_spin(const unsigned* p, unsigned target) {
  unsigned v;
  const long long t0 = clock64();
  while (true) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v >= target) break;
    if (clock64() - t0 > 4000000000LL) mbar_timeout_trap();
  }
}

template <bool kTf32, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1) conv_chain_kernel(const __grid_constant__ ChainParams P, const ConvMaps* __restrict__ maps, unsigned* done) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_slot_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot_s;
  const int G = gridDim.x;
  bool first = true;
  for (int l = 0; l < P.n_ops; ++l) {
    const ChainMeta& m = P.meta[l];
    int it = static_cast<int>(blockIdx.x) - m.cta_off;
    if (it < 0) it += G;
    for (; it < m.n_items; it += G) {
      long long* idbg = P.dbg ? P.dbg + (static_cast<long long>(l) * G + blockIdx.x) * 16 : nullptr;
      auto gstamp = [&](int slot) { if (idbg) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); idbg[slot] = static_cast<long long>(t); } };
      auto dep_wait = [&] {
        if (threadIdx.x == 64) {
          for (int k = 0; k < m.n_deps; ++k) chain_spin(done + m.dep[k], static_cast<unsigned>(m.dep_target[k]));
          asm volatile("fence.proxy.async;" ::: "memory");   // the producers wrote through TMA; this CTA reads through TMA
          gstamp(6);
        }
      };
      if (m.type == 0) {
        const ConvArgs& a = P.args[l];
        const int bx = it % m.gx, t2 = it / m.gx, by = t2 % m.gy, bz = t2 / m.gy;
        conv_tile<YP_FMT_F32X2, 2, kTf32, NT, true>(maps[l], a, bx, by, bz, m.gx, m.gy, smem_raw, tmem_base, first, dep_wait, idbg, m.variant);
        first = false;
      } else {
        const ChainPool& pp = P.pool[m.variant];
        dep_wait();
        __syncthreads();
        sppf_pool_item(pp.cat4, pp.C, it % pp.groups, it / pp.groups, smem_raw);   // ends with __syncthreads()
      }
      // every global write of the item has been performed (TMA stores: wait_group 0 by their issuing thread before the item's last
      // barrier; generic stores of the other threads: ordered before this release by that barrier)
      if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(done + l) : "memory");
        gstamp(7);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned old = atomicAdd(done + kMaxChain, 1u);
    if (old == static_cast<unsigned>(G - 1)) {   // every CTA has finished all of its items: nobody reads the counters any more
      for (int l = 0; l <= kMaxChain; ++l) done[l] = 0;
      __threadfence();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
long long* g_timeline = nullptr;
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  return g_encode;
}

CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// 5-D map over (C, W, H, B, plane) of an NHWC view; `sub` = spatial subsampling (2 -> parity view ph,pw).
int encode_view(CUtensorMap* tm, const YpView& v, int sub, int ph, int pw, int Wfull, int Hfull, int box_c, int box_w,
                int box_h, int box_p = 1) {
  const int es = fmt_esize(v.format);
  const int planes = fmt_planes(v.format);
  const CUtensorMapDataType dt = v.format == YP_FMT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  char* base = static_cast<char*>(v.base) + (static_cast<int64_t>(ph) * Wfull + pw) * v.pix_stride * es;
  cuuint64_t dims[5] = {(cuuint64_t)v.C, (cuuint64_t)(Wfull / sub), (cuuint64_t)(Hfull / sub), (cuuint64_t)v.B, (cuuint64_t)planes};
  const int64_t img = static_cast<int64_t>(Hfull) * Wfull * v.pix_stride;
  cuuint64_t strides[4] = {(cuuint64_t)(v.pix_stride * sub * es), (cuuint64_t)(v.pix_stride * Wfull * sub * es), (cuuint64_t)(img * es),
                           (cuuint64_t)((planes > 1 ? v.plane_stride : img * v.B) * es)};
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, (cuuint32_t)box_p};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  YP_REQUIRE(aligned16(base), YP_ERR_ALIGN, "conv: view base %p not 16-byte aligned", (void*)base);
  for (int i = 0; i < 4; ++i) YP_REQUIRE(strides[i] % 16 == 0, YP_ERR_ALIGN, "conv: view stride %d (%llu B) not a multiple of 16", i, (unsigned long long)strides[i]);
  CUresult r = get_encode()(tm, dt, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(box_c * es),
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YP_REQUIRE(r == CUDA_SUCCESS, YP_ERR_CUDA, "cuTensorMapEncodeTiled(view) failed: %d (C=%d W=%d H=%d B=%d box %d,%d,%d)", (int)r, v.C,
             Wfull / sub, Hfull / sub, v.B, box_c, box_w, box_h);
  return YP_OK;
}

int encode_weight(CUtensorMap* tm, const void* w, int fmt, int Ktot, int cout, int box_k, int box_n) {
  const int es = fmt_esize(fmt);
  const int planes = fmt_planes(fmt);
  const CUtensorMapDataType dt = fmt == YP_FMT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t dims[3] = {(cuuint64_t)Ktot, (cuuint64_t)cout, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)Ktot * es, (cuuint64_t)Ktot * cout * es};
  cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_n, (cuuint32_t)planes};  // one TMA brings both weight planes
  cuuint32_t estr[3] = {1, 1, 1};
  YP_REQUIRE(aligned16(w) && strides[0] % 16 == 0, YP_ERR_ALIGN, "conv: weight pointer/row stride not 16-byte aligned");
  CUresult r = get_encode()(tm, dt, 3, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            swizzle_for(box_k * es), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YP_REQUIRE(r == CUDA_SUCCESS, YP_ERR_CUDA, "cuTensorMapEncodeTiled(weight) failed: %d (K=%d N=%d box %d,%d)", (int)r, Ktot, cout, box_k, box_n);
  return YP_OK;
}

template <int OUT_FMT, int UNITS, bool kTf32, int NT>
int launch(const ConvMaps& maps, const ConvArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  auto kern = conv_tc_kernel<OUT_FMT, UNITS, kTf32, NT>;
  static thread_local size_t configured = 0;  // per (instantiation, thread): raise the dynamic smem limit once per size
  if (smem > configured) {
    YP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  static const bool use_pdl = getenv("YP_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
  YP_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps, a));
  return YP_OK;
}

template <int OUT_FMT, int UNITS, bool kTf32>
int launch_persist(const ConvMaps& maps, const ConvArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  auto kern = conv_tc_persist_kernel<OUT_FMT, UNITS, kTf32>;
  static thread_local bool configured = false;
  if (!configured) {
    YP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  static const bool use_pdl = getenv("YP_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kPersistThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
  YP_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps, a));
  return YP_OK;
}

template <int OUT_FMT>
int launch_drain(const ConvMaps& maps, const ConvArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  auto kern = conv_tc_drain_kernel<OUT_FMT>;
  static thread_local bool configured = false;
  if (!configured) {
    YP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  static const bool use_pdl = getenv("YP_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kPersistThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
  YP_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps, a));
  return YP_OK;
}

}  // namespace

void set_conv_timeline(long long* p) { g_timeline = p; }

// Tile-shape selection (fewest 128-row tiles; ties -> widest patch).
void pick_patch(int Ho, int Wo, int* Ht, int* Wt) {
  int best_tiles = 1 << 30, bw = 1, bh = 1;
  for (int wt = 1; wt <= Wo && wt <= 128; ++wt) {
    int ht = 128 / wt;
    if (ht > Ho) ht = Ho;
    if (ht < 1) continue;
    const int tiles = ceil_div(Wo, wt) * ceil_div(Ho, ht);
    const bool al = (wt * ht) % 8 == 0, bal = (bw * bh) % 8 == 0;  // 8-row aligned patches let one TMA bring both operand planes
    if (tiles < best_tiles || (tiles == best_tiles && ((al && !bal) || (al == bal && wt > bw)))) { best_tiles = tiles; bw = wt; bh = ht; }
  }
  *Ht = bh; *Wt = bw;
}

// Patch mode: M rows index an Ht x (Wt+2) padded patch, so Ht * (Wt + 2) <= 128.
void pick_patch_padded(int Ho, int Wo, int* Ht, int* Wt) {
  int best_tiles = 1 << 30, bw = 1, bh = 1;
  for (int wt = 1; wt <= Wo && wt <= 126; ++wt) {
    int ht = 128 / (wt + 2);
    if (ht > Ho) ht = Ho;
    if (ht < 1) continue;
    const int tiles = ceil_div(Wo, wt) * ceil_div(Ho, ht);
    if (tiles < best_tiles || (tiles == best_tiles && wt > bw)) { best_tiles = tiles; bw = wt; bh = ht; }
  }
  *Ht = bh; *Wt = bw;
}

namespace {

// > 0 while plan_conv() plans an operation of a layer chain (conv_chain_kernel): shared-memory budget in bytes for the pipeline /
// staging region; the plan is then the one-CTA-per-SM kind (up to 512 TMEM columns, no persistent variant, kChainThreads threads).
thread_local int g_chain_budget = 0;
constexpr int kChainThreads = 256;

struct ConvPlan {
  ConvArgs a;
  dim3 grid;
  size_t smem, ws_bytes, ws_counter_bytes;
  int out_fmt, units, chunk_elems;
  bool tf32;
  int nt;   // threads per CTA: 256 (two CTAs per SM) or 512 (one CTA per SM, four epilogue warps per TMEM lane quarter)
};

// Everything that does not need device pointers: tile shapes, accumulator plan, split-K factor, workspace need.
int plan_conv_impl(const YpConvDesc& d, ConvPlan* P, bool allow_split, bool no_patch, bool* patch_no_fit) {
  const YpView& in = d.in;
  const int in_fmt = in.format;
  YP_REQUIRE(in_fmt == YP_FMT_F32X2 || in_fmt == YP_FMT_BF16, YP_ERR_SHAPE, "conv: input format %d unsupported", in_fmt);
  const bool tf32 = in_fmt == YP_FMT_F32X2;
  const int es = fmt_esize(in_fmt);
  YP_REQUIRE((d.ksize == 1 && d.stride == 1) || (d.ksize == 3 && (d.stride == 1 || d.stride == 2)) ||
                 (d.ksize == 0 && d.stride == 1 && d.n_taps >= 1 && d.n_taps <= 9), YP_ERR_SHAPE,
             "conv: k=%d s=%d (n_taps=%d) unsupported", d.ksize, d.stride, d.n_taps);
  YP_REQUIRE(in.C % 16 == 0 && d.cout % 16 == 0, YP_ERR_SHAPE, "conv: Cin=%d / Cout=%d must be multiples of 16", in.C, d.cout);
  YP_REQUIRE(d.stride == 1 || (in.H % 2 == 0 && in.W % 2 == 0), YP_ERR_SHAPE, "conv: stride 2 needs even H,W");
  YP_REQUIRE((d.n_out >= 1 && d.n_out <= 2) || ((d.epilogue & YP_EPI_ROWMIN) && d.n_out == 0), YP_ERR_SHAPE, "conv: n_out=%d", d.n_out);
  const int Ho = in.H / d.stride, Wo = in.W / d.stride;

  ConvArgs& a = P->a;
  memset(&a, 0, sizeof(a));
  P->tf32 = tf32;
  a.Ho = Ho; a.Wo = Wo; a.ksize = d.ksize; a.stride = d.stride;
  pick_patch(Ho, Wo, &a.Ht, &a.Wt);
  a.tiles_w = ceil_div(Wo, a.Wt); a.tiles_h = ceil_div(Ho, a.Ht);
  a.n_taps = d.ksize ? d.ksize * d.ksize : d.n_taps;
  for (int t = 0; t < a.n_taps && d.ksize == 0; ++t) {
    YP_REQUIRE(d.tap_dh[t] >= -8 && d.tap_dh[t] <= 7 && d.tap_dw[t] >= -8 && d.tap_dw[t] <= 7, YP_ERR_SHAPE, "conv: tap offset out of range");
    a.tap_dh |= static_cast<unsigned long long>(d.tap_dh[t] + 8) << (4 * t);
    a.tap_dw |= static_cast<unsigned long long>(d.tap_dw[t] + 8) << (4 * t);
  }
  const int cin_bytes = in.C * es;
  a.ck_bytes = cin_bytes % 128 == 0 ? 128 : (cin_bytes % 64 == 0 ? 64 : 32);
  a.ck_elems = a.ck_bytes / es;
  a.kb_per_tap = in.C / a.ck_elems;
  a.in_planes = tf32 ? 2 : 1;
  static const bool allow_patch = getenv("YP_CONV_NO_PATCH") == nullptr;
  static const int base_offset_mode = getenv("YP_CONV_BASE_OFFSET") ? atoi(getenv("YP_CONV_BASE_OFFSET")) : 0;
  a.patch = (allow_patch && !no_patch && d.ksize == 3 && d.stride == 1 && a.ck_bytes >= 64 && !(d.epilogue & YP_EPI_NO_PATCH)) ? 1 : 0;
  a.base_offset_mode = base_offset_mode;
  if (a.patch) {
    pick_patch_padded(Ho, Wo, &a.Ht, &a.Wt);
    a.tiles_w = ceil_div(Wo, a.Wt); a.tiles_h = ceil_div(Ho, a.Ht);
    a.Wp = a.Wt + 2;
    a.a_rows = (a.Ht + 2) * a.Wp;
  }
  const int rows = a.patch ? a.a_rows : a.Ht * a.Wt;           // rows one TMA box of the A operand writes
  // rows the tensor core may touch: the last tap's window starts (2*Wp+2) rows into the patch and always spans 128 rows
  const int rows_alloc = a.patch ? ((std::max(a.a_rows, 128 + 2 * a.Wp + 2) + 7) & ~7) : 128;
  a.a_split = (!a.patch && a.in_planes == 2 && rows % 8 == 0) ? 1 : 0;   // lo plane lands right behind the hi plane, swizzle-aligned
  a.a_plane_off = a.patch ? rows_alloc * a.ck_bytes : ((rows + 7) & ~7) * a.ck_bytes;
  // K loop units: (tap, channel block) pairs, or channel blocks (each covering all nine taps) in patch mode
  const int num_kb = a.patch ? a.kb_per_tap : a.n_taps * a.kb_per_tap;
  const int ksteps = (a.patch ? 9 : 1) * (a.ck_bytes / 32);   // MMA k-steps per K-loop unit

  // ---- output format / staging geometry
  const bool rowmin = (d.epilogue & YP_EPI_ROWMIN) != 0;
  const int out_fmt = rowmin ? YP_FMT_F32 : d.out[0].format;
  P->out_fmt = out_fmt;
  for (int i = 0; i < d.n_out; ++i) {
    YP_REQUIRE(d.out[i].format == out_fmt, YP_ERR_SHAPE, "conv: all outputs must share one format");
    YP_REQUIRE(d.out[i].C == d.cout && d.out[i].B == in.B && d.out[i].H == Ho && d.out[i].W == Wo, YP_ERR_SHAPE,
               "conv: output %d geometry mismatch (C %d vs %d, HxW %dx%d vs %dx%d)", i, d.out[i].C, d.cout, d.out[i].H, d.out[i].W, Ho, Wo);
  }
  YP_REQUIRE(tf32 ? (out_fmt == YP_FMT_F32X2 || out_fmt == YP_FMT_F32) : (out_fmt == YP_FMT_BF16 || out_fmt == YP_FMT_F32), YP_ERR_SHAPE,
             "conv: output format %d incompatible with input format %d", out_fmt, in_fmt);
  const int oes = fmt_esize(out_fmt);
  a.out_planes = fmt_planes(out_fmt);
  a.out_fmt = out_fmt;
  const int cout_bytes = d.cout * oes;
  a.out_row_bytes = cout_bytes % 128 == 0 ? 128 : (cout_bytes % 64 == 0 ? 64 : 32);
  const int chunk_elems = a.out_row_bytes / oes;
  YP_REQUIRE(chunk_elems % 16 == 0, YP_ERR_SHAPE, "conv: Cout=%d gives a %d-element store chunk (<16)", d.cout, chunk_elems);
  P->chunk_elems = chunk_elems;
  P->units = chunk_elems / 16;
  a.staging_set_bytes = a.out_planes * 128 * a.out_row_bytes;

  // ---- N tile: the largest divisor of Cout (<= 128 in 3xTF32 mode so that [W_hi; W_lo] stacks into one N <= 256 MMA) that
  // still yields >= #SM CTAs; when the layer cannot fill the GPU anyway, the smallest tile >= 32 (latency).
  const int m_tiles = a.tiles_w * a.tiles_h * in.B;
  const int nsm = sm_count();
  int Nt = 0;
  int split_req = d.split_k;           // 0 = heuristic, 1 = never, n = n slices
  // Throughput plan on the drained-accumulator kernel (YP_CONV_DRAIN=0 switches it off): 3xTF32, 128-byte store chunks, no L2-norm /
  // row-min epilogue, no custom tap list, not inside a layer chain.
  static const bool allow_drain = getenv("YP_CONV_DRAIN") == nullptr || atoi(getenv("YP_CONV_DRAIN")) != 0;
  // Layers with a short K (few MMAs per accumulator anyway) keep the rotating-accumulator plans below, which run two CTAs per SM.
  static const int drain_min_steps = getenv("YP_CONV_DRAIN_MIN_STEPS") ? atoi(getenv("YP_CONV_DRAIN_MIN_STEPS")) : 0;
  const bool drain = allow_drain && d.tile_n <= YP_TILE_WIDE && tf32 && g_chain_budget == 0 && d.ksize != 0 && chunk_elems == 32 && d.cout % 32 == 0 &&
                     !(d.epilogue & (YP_EPI_L2NORM | YP_EPI_ROWMIN)) && (out_fmt == YP_FMT_F32X2 || out_fmt == YP_FMT_F32) &&
                     num_kb * ksteps >= drain_min_steps;
  if (d.epilogue & YP_EPI_L2NORM) {
    YP_REQUIRE(d.cout <= 256, YP_ERR_SHAPE, "conv: L2-norm epilogue needs Cout <= 256 (got %d)", d.cout);
    Nt = d.cout;
  } else if (drain) {
    // drained accumulators (conv_tc_drain_kernel): chains are bounded at any width -> the widest tile, K never split
    for (int n = 128; n >= 32; n -= 32)
      if (d.cout % n == 0) { Nt = n; break; }
    split_req = 1;
  } else if (d.tile_n <= YP_TILE_WIDE) {
    // Throughput plan: the widest N tile (fewest re-reads of the activation tile, fewest MMAs per FLOP) whose accumulator plan still
    // keeps the fp32-grade accuracy of the 3xTF32 mode.  The tensor core adds into an fp32 accumulator with truncation, so the error
    // grows with the number of MMAs chained on one accumulator: a wide tile leaves TMEM room for fewer accumulators to rotate over
    // (Nt = 128: three main accumulators beside one for both cross terms -- the un-stacked plan below --, 64: three [main | cross]
    // pairs, 32: six).  Every candidate tile is charged the K split that keeps its chains <= kMaxChain main MMAs (each slice has its
    // own accumulators; partial tiles are summed in fp32 with rounding).
    static const int kMaxChain = getenv("YP_CONV_MAX_CHAIN") ? atoi(getenv("YP_CONV_MAX_CHAIN")) : 24;
    const int total_steps = num_kb * ksteps;
    auto n_main = [&](int nt) {
      if (!tf32) return 1 << 20;       // bf16 operands: the accumulator plan is not the accuracy limit
      if (nt == 128 && m_tiles < nsm) return 3;     // un-stacked plan (single-wave layers only: it needs all 512 TMEM columns, one CTA per SM)
      const int lim = (static_cast<long long>(m_tiles) * (d.cout / nt) > nsm) ? 256 : 512;
      int n_s = 2, n_p = (lim - n_s * nt) / (2 * nt);
      if (n_p < 1) { n_s = 1; n_p = (lim - nt) / (2 * nt); }
      if (n_p < 1) n_p = (512 - nt) / (2 * nt);
      return std::max(1, std::min(n_p, 6));
    };
    // Candidates from wide to narrow; the first that needs no K split wins.  When every tile needs a split: a layer that fills the GPU
    // anyway takes the tile with the fewest slices (split-K partial tiles are pure overhead there), a small layer the widest tile
    // (its extra CTAs only cost their fixed set-up, the tensor time per FLOP is what the wide tile saves).
    int best_split = 1 << 20, widest = 0, widest_split = 1;
    for (int n = tf32 ? 128 : 256; n >= chunk_elems; n -= 16) {
      if (d.cout % n || n % chunk_elems) continue;
      int need = ceil_div(total_steps, n_main(n) * kMaxChain);          // K slices that bound the chains of this tile
      if (need > num_kb) need = num_kb;
      if (!widest) { widest = n; widest_split = need; }
      if (need < best_split) { Nt = n; best_split = need; }
      if (n <= 32 || best_split == 1) break;
    }
    if (best_split > 1 && static_cast<long long>(m_tiles) * (d.cout / widest) < nsm / 2) { Nt = widest; best_split = widest_split; }
    split_req = std::max(1, best_split);
  } else if (d.tile_n > 0) {
    YP_REQUIRE(d.tile_n % chunk_elems == 0 && d.cout % d.tile_n == 0 && d.tile_n <= (tf32 ? 128 : 256), YP_ERR_SHAPE,
               "conv: tile_n=%d invalid for Cout=%d (store chunk %d)", d.tile_n, d.cout, chunk_elems);
    Nt = d.tile_n;
  } else {
    const int nmax = tf32 ? 128 : 256;
    int smallest = 0;
    for (int n = nmax; n >= chunk_elems; n -= 16) {
      if (d.cout % n || n % chunk_elems) continue;
      smallest = n;
      if (Nt == 0 && m_tiles * (d.cout / n) >= nsm / 2) Nt = n;   // measured: wider tiles win as soon as half the SMs are busy
      if (n <= 32 && smallest) break;
    }
    if (Nt == 0) Nt = smallest;
  }
  YP_REQUIRE(Nt >= 16 && Nt % 16 == 0 && Nt <= 256, YP_ERR_SHAPE, "conv: no valid N tile for Cout=%d", d.cout);
  a.Nt = Nt;
  const int n_tiles = d.cout / Nt;

  // ---- split-K: deep layers on small feature maps occupy few CTAs; slice K over grid.z so that ~#SM CTAs are busy.
  int S = 1;
  if (allow_split && !rowmin && split_req != 1 && static_cast<size_t>(m_tiles) * n_tiles * sizeof(int) <= kWsCounterBytes) {
    const int ctas = m_tiles * n_tiles;
    int want = split_req > 1 ? split_req : nsm / ctas;
    if (want > 16) want = 16;
    if (split_req <= 1) {                         // heuristic: at most 8 slices of >= 4 k-blocks (>= 1 channel block in patch mode)
      if (want > 8) want = 8;
      const int lim = a.patch ? num_kb : num_kb / 4;
      if (want > lim) want = lim;
    } else if (want > num_kb) want = num_kb;
    if (want >= 2) S = want;
  }
  a.kb_per_split = ceil_div(num_kb, S);
  S = ceil_div(num_kb, a.kb_per_split);
  a.split_k = S;
  const int kb_min = num_kb - (S - 1) * a.kb_per_split;   // k-blocks of the last (shortest) slice
  const int mmas_min = kb_min * ksteps;                    // main MMAs every CTA issues at least

  // Layers with more CTAs than SMs run two CTAs per SM (the epilogue of one overlaps the main loop of the other): each CTA
  // then gets half of the shared memory and at most 256 TMEM columns.
  static const bool allow_dense = getenv("YP_CONV_NO_DENSE") == nullptr;
  // Several waves of tiles in bf16: the persistent kernel (one CTA per SM, double-buffered TMEM accumulators) instead of two
  // independent CTAs per SM.  YP_CONV_PERSIST=0 switches it off.
  static const bool allow_persist = getenv("YP_CONV_PERSIST") == nullptr || atoi(getenv("YP_CONV_PERSIST")) != 0;
  const bool chain = g_chain_budget > 0;
  const bool persist = allow_persist && !chain && !tf32 && S == 1 && !(d.epilogue & (YP_EPI_L2NORM | YP_EPI_ROWMIN)) && out_fmt != YP_FMT_F32X2 &&
                       static_cast<long long>(m_tiles) * n_tiles > (getenv("YP_CONV_PERSIST_MIN") ? atoll(getenv("YP_CONV_PERSIST_MIN")) : 2LL * nsm);
  static const bool force_dense = getenv("YP_CONV_FORCE_DENSE") != nullptr && atoi(getenv("YP_CONV_FORCE_DENSE")) != 0;
  const bool unstacked = !drain && d.tile_n <= YP_TILE_WIDE && m_tiles < nsm && tf32 && Nt == 128 && !(d.epilogue & YP_EPI_L2NORM);   // needs all 512 TMEM columns
  const bool dense = !chain && !unstacked && !drain && ((allow_dense && (force_dense || static_cast<long long>(m_tiles) * n_tiles * S > nsm)) || persist);   // persist: 256 columns per accumulator buffer
  const int tmem_limit = dense ? 256 : 512;

  // ---- accumulator / issuer plan (see the kernel comment)
  const uint32_t ab = tf32 ? 2u : 1u;
  auto idesc = [&](int n) {  // cute::UMMA::InstrDescriptor: c F32 [4,6)=1, a/b format [7,10)/[10,13), K-major, N>>3 [17,23), M>>4 [24,29)
    return (1u << 4) | (ab << 7) | (ab << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
  };
  int cols = 0;
  if (drain) {
    // conv_tc_drain_kernel: main product in two round buffers, both cross terms in one accumulator per tile (two tile buffers)
    a.n_iss = 2;
    static const bool drain_stack = getenv("YP_CONV_DRAIN_STACK") == nullptr || atoi(getenv("YP_CONV_DRAIN_STACK")) != 0;
    a.buf_cols = (drain_stack && Nt <= 64) ? 2 * Nt : Nt;      // stacked round buffers need 6 Nt <= 512 TMEM columns
    a.iss_idesc[0] = idesc(a.buf_cols);
    a.iss_idesc[1] = idesc(Nt);
    a.n_src = 0;
    cols = 2 * a.buf_cols + 2 * Nt;
    static const int drain_steps = getenv("YP_CONV_DRAIN_STEPS") ? atoi(getenv("YP_CONV_DRAIN_STEPS")) : 4;   // MMAs chained per round
    a.drain_units = std::max(1, drain_steps / (a.ck_bytes / 32));
  } else if (unstacked) {
    // Wide throughput tile: the main product A_hi x W_hi rotates over three accumulators (issuer 0), both cross terms go into a fourth
    // (issuer 1, two MMAs per k-step; their partial sums are ~2^-11 of the result, so their truncation error is negligible).
    const int n_p = std::min(3, std::max(1, mmas_min));
    a.n_iss = 2; a.kstart[0] = a.kstart[1] = 0; a.kinc[0] = a.kinc[1] = 1;
    a.n_jobs[0] = 1; a.job_a[0][0] = 0; a.job_b[0][0] = 0; a.iss_col[0] = 0; a.iss_stride[0] = Nt; a.iss_cnt[0] = n_p; a.iss_idesc[0] = idesc(Nt);
    a.n_jobs[1] = 2; a.job_a[1][0] = 1; a.job_b[1][0] = 0; a.job_a[1][1] = 0; a.job_b[1][1] = 1;
    a.iss_col[1] = n_p * Nt; a.iss_stride[1] = 0; a.iss_cnt[1] = 1; a.iss_idesc[1] = idesc(Nt);
    for (int j = 0; j < n_p; ++j) a.src_col[j] = j * Nt;
    a.src_col[n_p] = n_p * Nt;
    a.n_src = n_p + 1;
    cols = (n_p + 1) * Nt;
  } else if (tf32 && Nt <= 128) {
    // issuer 0: A_hi x [W_hi; W_lo]  (N = 2 Nt)  -> pairs [main | cross2];  issuer 1: A_lo x W_hi (N = Nt) -> cross1
    int n_s = 2, n_p = (tmem_limit - n_s * Nt) / (2 * Nt);
    if (n_p < 1) { n_s = 1; n_p = (tmem_limit - Nt) / (2 * Nt); }
    if (n_p < 1) { n_s = 1; n_p = (512 - Nt) / (2 * Nt); }
    if (n_p > 6) n_p = 6;
    if (n_p > mmas_min) n_p = mmas_min;
    if (n_s > mmas_min) n_s = mmas_min;
    YP_REQUIRE(n_p >= 1 && n_s >= 1, YP_ERR_SHAPE, "conv: no accumulator plan for Nt=%d", Nt);
    // Issue rate: one thread sustains one small MMA per ~120 cycles (descriptor / accumulator bookkeeping), the tensor core needs
    // 2 Nt / 2 + Nt / 2 cycles for the two MMAs of a k-step (48 for Nt = 32), so each of the two MMA streams is dealt to TWO
    // threads (even / odd k-steps of every k-block, each with its own accumulators) when it has the accumulators and k-steps.
    static const bool four = getenv("YP_CONV_ISSUERS4") == nullptr || atoi(getenv("YP_CONV_ISSUERS4")) != 0;
    const int per_blk = a.ck_bytes / 32;
    const bool split_main = four && per_blk >= 2 && n_p >= 2 && mmas_min >= 2 * n_p;
    const bool split_cross = four && per_blk >= 2 && n_s >= 2 && mmas_min >= 2 * n_s;
    int qn = 0;
    auto add_iss = [&](int ja, int col, int stride, int cnt, int n, int ks, int ki) {
      a.n_jobs[qn] = 1; a.job_a[qn][0] = ja; a.job_b[qn][0] = 0; a.iss_col[qn] = col; a.iss_stride[qn] = stride; a.iss_cnt[qn] = cnt;
      a.iss_idesc[qn] = idesc(n); a.kstart[qn] = ks; a.kinc[qn] = ki;
      ++qn;
    };
    if (split_main) {
      const int h0n = (n_p + 1) / 2;
      add_iss(0, 0, 2 * Nt, h0n, 2 * Nt, 0, 2);
      add_iss(0, h0n * 2 * Nt, 2 * Nt, n_p - h0n, 2 * Nt, 1, 2);
    } else {
      add_iss(0, 0, 2 * Nt, n_p, 2 * Nt, 0, 1);
    }
    if (split_cross) {
      add_iss(1, n_p * 2 * Nt, Nt, 1, Nt, 0, 2);
      add_iss(1, n_p * 2 * Nt + Nt, Nt, n_s - 1, Nt, 1, 2);
    } else {
      add_iss(1, n_p * 2 * Nt, Nt, n_s, Nt, 0, 1);
    }
    a.n_iss = qn;
    int k = 0;
    for (int j = 0; j < n_p; ++j) a.src_col[k++] = j * 2 * Nt;               // main products first
    for (int j = 0; j < n_p; ++j) a.src_col[k++] = j * 2 * Nt + Nt;          // A_hi x W_lo
    for (int j = 0; j < n_s; ++j) a.src_col[k++] = n_p * 2 * Nt + j * Nt;    // A_lo x W_hi
    a.n_src = k;
    cols = n_p * 2 * Nt + n_s * Nt;
  } else if (tf32) {
    // Nt in (128, 256] (L2-norm head of wide models): main accumulator + one accumulator for both cross terms
    a.n_iss = 2; a.kstart[0] = a.kstart[1] = 0; a.kinc[0] = a.kinc[1] = 1;
    a.n_jobs[0] = 1; a.job_a[0][0] = 0; a.job_b[0][0] = 0; a.iss_col[0] = 0; a.iss_stride[0] = 0; a.iss_cnt[0] = 1; a.iss_idesc[0] = idesc(Nt);
    a.n_jobs[1] = 2; a.job_a[1][0] = 1; a.job_b[1][0] = 0; a.job_a[1][1] = 0; a.job_b[1][1] = 1;
    a.iss_col[1] = Nt; a.iss_stride[1] = 0; a.iss_cnt[1] = 1; a.iss_idesc[1] = idesc(Nt);
    a.n_src = 2; a.src_col[0] = 0; a.src_col[1] = Nt;
    cols = 2 * Nt;
  } else {
    const int total = (tmem_limit / Nt) >= 1 ? tmem_limit / Nt : 512 / Nt;
    static const int max_iss = getenv("YP_CONV_ISSUERS") ? atoi(getenv("YP_CONV_ISSUERS")) : 2;
    a.n_iss = (total >= 2 && ksteps >= 2 && max_iss >= 2) ? 2 : 1;
    // one accumulator per issuer: every accumulator costs a 128-lane TMEM read in the epilogue; short K loops (<= 16
    // k-steps, the 1x1 layers) are issued by a single thread into a single accumulator
    if (mmas_min <= 16) a.n_iss = 1;
    for (int q = 0; q < 2; ++q) { a.kstart[q] = a.n_iss == 2 ? q : 0; a.kinc[q] = a.n_iss == 2 ? 2 : 1; }
    int each = 1;
    const int per = mmas_min / a.n_iss;
    if (each > per) each = per;
    YP_REQUIRE(each >= 1, YP_ERR_SHAPE, "conv: no accumulator plan for Nt=%d", Nt);
    int k = 0;
    for (int q = 0; q < a.n_iss; ++q) {
      a.n_jobs[q] = 1; a.job_a[q][0] = 0; a.job_b[q][0] = 0; a.iss_col[q] = q * each * Nt; a.iss_stride[q] = Nt; a.iss_cnt[q] = each; a.iss_idesc[q] = idesc(Nt);
      for (int j = 0; j < each; ++j) a.src_col[k++] = (q * each + j) * Nt;
    }
    a.n_src = k;
    cols = a.n_iss * each * Nt;
  }
  YP_REQUIRE(cols <= 512 && a.n_src <= 16, YP_ERR_SHAPE, "conv: accumulator plan needs %d TMEM columns", cols);
  a.tmem_cols = 32;
  while ((int)a.tmem_cols < cols) a.tmem_cols <<= 1;

  // ---- pipeline geometry
  a.a_region_bytes = a.in_planes * 128 * a.ck_bytes;
  const int b_region_bytes = a.in_planes * Nt * a.ck_bytes;
  a.stage_bytes = a.a_region_bytes + b_region_bytes;
  // TMA counts the bytes of the boxes actually written: Ht*Wt (<= 128) rows per A plane, Nt rows per B plane
  a.tx_bytes = a.in_planes * (rows * a.ck_bytes + Nt * a.ck_bytes);
  int budget = chain ? g_chain_budget : ((persist || drain) ? 198 * 1024 - 2 * a.staging_set_bytes : (dense ? 104 * 1024 : 200 * 1024));
  int region = 0;
  if (a.patch) {
    a.a_stage_bytes = a.in_planes * rows_alloc * a.ck_bytes;
    a.a_tx = a.in_planes * a.a_rows * a.ck_bytes;
    a.b_stage_bytes = b_region_bytes;
    a.b_tx = b_region_bytes;
    a.b_ring_off = 2 * a.a_stage_bytes;
    if (2 * a.a_stage_bytes + 2 * a.b_stage_bytes > budget) {
      if (drain) { *patch_no_fit = true; return YP_ERR_SHAPE; }   // the staging sets live beside the pipeline: re-plan with per-tap loads
      budget = 200 * 1024;
    }
    int bs = (budget - 2 * a.a_stage_bytes) / a.b_stage_bytes;
    if (bs > kMaxStages) bs = kMaxStages;
    if (bs < 2) { *patch_no_fit = true; return YP_ERR_SHAPE; }   // caller re-plans with per-tap loads
    a.b_stages = bs;
    a.stages = bs;
    region = 2 * a.a_stage_bytes + bs * a.b_stage_bytes;
  } else {
    if (a.stage_bytes * 2 > budget && !drain) budget = 200 * 1024;   // keep at least two stages
    int stages = budget / a.stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > a.kb_per_split) stages = a.kb_per_split;
    if (stages < 1) stages = 1;
    YP_REQUIRE(a.stage_bytes <= 200 * 1024, YP_ERR_SHAPE, "conv: stage of %d bytes exceeds shared memory", a.stage_bytes);
    a.stages = stages;
    region = stages * a.stage_bytes;
  }
  // Layers that cannot fill the GPU (at most one CTA per SM) run 512-thread CTAs: 16 epilogue warps instead of 8.
  // Measured on B200 (YOLOPoint-S 640x640 batch 1, fp32): 0.786 ms per network pass with it, 0.777 ms without -- a 512-thread CTA owns the
  // whole register file of its SM, so the next layer's CTAs can no longer start under this layer's tail (programmatic dependent launch).
  // Off unless YP_CONV_WIDE=1.
  static const bool allow_wide = getenv("YP_CONV_WIDE") != nullptr && atoi(getenv("YP_CONV_WIDE")) != 0;
  P->nt = chain ? kChainThreads : (drain ? kPersistThreads : ((allow_wide && !dense && !persist) ? 512 : 256));
  const int n_groups = P->nt / 128;
  const int staging_sets = 2 * (n_groups > P->units ? n_groups / P->units : 1);
  if (persist || drain) {
    // the staging sets may not alias the pipeline stages: the next tile's K loop runs under this tile's epilogue
    region = (region + 1023) & ~1023;
    a.stg_off = region;
    region += 2 * a.staging_set_bytes;
  } else if (region < staging_sets * a.staging_set_bytes) region = staging_sets * a.staging_set_bytes;
  region = (region + 1023) & ~1023;
  a.bar_off = region;
  P->smem = 1024 /*alignment slack*/ + region + 8 * (2 * kMaxStages + 14) + Nt * sizeof(float) + 16;
  YP_REQUIRE(P->smem <= 227 * 1024, YP_ERR_SHAPE, "conv: needs %zu bytes of shared memory", P->smem);
  P->grid = dim3(m_tiles, n_tiles, S);
  a.persist = persist ? 1 : (drain ? 2 : 0);
  if (drain) {
    a.n_tiles_n = n_tiles;
    a.tiles_total = m_tiles * n_tiles;
    // Persistent grid: at most nsm / drain_grid_div CTAs, every CTA an equal share of the tiles.  With several frames in flight a
    // layer does not need every SM, and a CTA that walks two or three tiles overlaps the epilogue of one with the main loop of the
    // next.  Measured (YOLOPoint-S 640x640 batch 1, 8 frames in flight): divisor 1 / 2 / 3 / 4 = 2372 / 2424 / 2461 / 2404 frames/s.
    // The divisor comes with the descriptor (tile_n = -g, YP_TILE_WIDE = -1 = every SM): the caller knows how many frames it keeps
    // in flight; YP_CONV_DRAIN_GRID_DIV overrides it.
    static const int env_div = getenv("YP_CONV_DRAIN_GRID_DIV") ? std::max(1, atoi(getenv("YP_CONV_DRAIN_GRID_DIV"))) : 0;
    const int grid_div = env_div ? env_div : std::max(1, -d.tile_n);
    const int max_ctas = std::max(1, nsm / grid_div);
    const int per_cta = ceil_div(a.tiles_total, std::min(a.tiles_total, max_ctas));
    P->grid = dim3(ceil_div(a.tiles_total, per_cta), 1, 1);
  }
  if (persist) {
    YP_REQUIRE(a.n_src <= 2 && cols <= 256, YP_ERR_SHAPE, "conv(persist): accumulator plan needs %d sources / %d columns", a.n_src, cols);
    a.n_tiles_n = n_tiles;
    a.tiles_total = m_tiles * n_tiles;
    a.buf_cols = 256;
    P->grid = dim3(std::min(a.tiles_total, nsm), 1, 1);
  }
  // The arrival counters live in a FIXED-size area at the start of the workspace: layers of one lane share the workspace,
  // and a per-layer counter area would overlap the partial sums an earlier (smaller) layer left behind.
  P->ws_counter_bytes = kWsCounterBytes;
  P->ws_bytes = S > 1 ? P->ws_counter_bytes + static_cast<size_t>(S) * m_tiles * n_tiles * 128 * Nt * sizeof(float) : 0;
  return YP_OK;
}

// Patch mode keeps two input patches resident; when the weight ring no longer fits beside them (wide L2-norm head)
// the layer is planned with per-tap loads instead.
int plan_conv(const YpConvDesc& d, ConvPlan* P, bool allow_split) {
  bool no_fit = false;
  const int rc = plan_conv_impl(d, P, allow_split, false, &no_fit);
  if (rc != YP_OK && no_fit) return plan_conv_impl(d, P, allow_split, true, &no_fit);
  return rc;
}

}  // namespace

size_t conv_tc_workspace_bytes(const YpConvDesc& d) {
  ConvPlan P;
  if (plan_conv(d, &P, true) != YP_OK) return 0;
  return P.ws_bytes;
}

// Shape / alignment / resource checks and tile planning only (host code, no CUDA calls): what yp_conv2d_plan_check returns.
int conv_tc_plan_check(const YpConvDesc& d) {
  ConvPlan P;
  return plan_conv(d, &P, true);
}

namespace {

// Runtime half of a planned launch: pointers of the descriptor into the kernel arguments, tensor maps of the operand / destination
// views.  `workspace` (split-K plans only): P.ws_bytes bytes, counters zero-initialised.
int prepare_conv(const YpConvDesc& d, ConvPlan& P, void* workspace, ConvMaps& maps) {
  int rc = YP_OK;
  ConvArgs& a = P.a;
  const YpView& in = d.in;
  const int Ho = a.Ho, Wo = a.Wo, Nt = a.Nt, out_fmt = P.out_fmt, chunk_elems = P.chunk_elems;
  if (a.split_k > 1) {
    YP_REQUIRE(aligned16(workspace), YP_ERR_ALIGN, "conv: workspace not 16-byte aligned");
    a.ws_counter = static_cast<int*>(workspace);
    a.ws_partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + P.ws_counter_bytes);
  }
  memset(&maps, 0, sizeof(maps));

  a.dbg = g_timeline;
  static const int ablate = getenv("YP_CONV_ABLATE") ? atoi(getenv("YP_CONV_ABLATE")) : 0;
  a.ablate = ablate;
  a.bias = d.bias;
  a.act = d.act;
  a.l2norm = (d.epilogue & YP_EPI_L2NORM) ? 1 : 0;
  a.rowmin = (d.epilogue & YP_EPI_ROWMIN) ? 1 : 0;
  a.row_key = d.row_key; a.col_key = d.col_key; a.n_rows = d.n_rows; a.n_cols = d.n_cols; a.col_off = d.col_off;
  if (a.rowmin && a.col_key) YP_REQUIRE(a.Ht == 1 && a.Wt == 128 && !a.patch, YP_ERR_SHAPE, "conv: YP_EPI_ROWMIN column keys expect 1 x 128 pixel tiles (got %d x %d)", a.Ht, a.Wt);
  if (a.rowmin) YP_REQUIRE(in.B == 1 && in.H == 1, YP_ERR_SHAPE, "conv: YP_EPI_ROWMIN expects the descriptors of set 1 as a [1, 1, N1, D] view");
  if (d.residual.base) {
    YP_REQUIRE(d.residual.C == d.cout && d.residual.H == Ho && d.residual.W == Wo && d.residual.B == in.B, YP_ERR_SHAPE, "conv: residual geometry mismatch");
    YP_REQUIRE(d.residual.format == out_fmt, YP_ERR_SHAPE, "conv: residual format %d must equal the output format %d", d.residual.format, out_fmt);
    YP_REQUIRE(aligned16(d.residual.base) && d.residual.pix_stride % 8 == 0 && d.residual.plane_stride % 4 == 0, YP_ERR_ALIGN, "conv: residual view not 16-byte aligned");
    a.res_base = d.residual.base; a.res_pix = d.residual.pix_stride; a.res_plane = d.residual.plane_stride;
  }

  // ---- tensor maps
  if (d.stride == 1) {
    if (a.patch) {
      if ((rc = encode_view(&maps.in[0], in, 1, 0, 0, in.W, in.H, a.ck_elems, a.Wp, a.Ht + 2, 1)) != YP_OK) return rc;
    } else if ((rc = encode_view(&maps.in[0], in, 1, 0, 0, in.W, in.H, a.ck_elems, a.Wt, a.Ht, a.a_split ? 2 : 1)) != YP_OK) return rc;
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw)
        if ((rc = encode_view(&maps.in[ph * 2 + pw], in, 2, ph, pw, in.W, in.H, a.ck_elems, a.Wt, a.Ht, a.a_split ? 2 : 1)) != YP_OK) return rc;
  }
  if ((rc = encode_weight(&maps.w, d.weight, in.format, a.n_taps * in.C, d.cout, a.ck_elems, Nt)) != YP_OK) return rc;
  int nm = 0;
  for (int i = 0; i < d.n_out; ++i) {
    const YpView& o = d.out[i];
    if (o.upsample == 2) {
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw)
          if ((rc = encode_view(&maps.out[nm++], o, 2, ph, pw, 2 * Wo, 2 * Ho, chunk_elems, a.Wt, a.Ht)) != YP_OK) return rc;
    } else if (o.upsample >= YP_UP_PARITY && o.upsample < YP_UP_PARITY + 4) {   // one parity class of a 2x larger destination
      const int pp = o.upsample - YP_UP_PARITY;
      if ((rc = encode_view(&maps.out[nm++], o, 2, pp >> 1, pp & 1, 2 * Wo, 2 * Ho, chunk_elems, a.Wt, a.Ht)) != YP_OK) return rc;
    } else {
      if ((rc = encode_view(&maps.out[nm++], o, 1, 0, 0, Wo, Ho, chunk_elems, a.Wt, a.Ht)) != YP_OK) return rc;
    }
  }
  a.n_out_maps = nm;
  return YP_OK;
}

}  // namespace

int conv_tc_forward(const YpConvDesc& d, cudaStream_t st) {
  YP_REQUIRE(get_encode() != nullptr, YP_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  ConvPlan P;
  int rc = plan_conv(d, &P, d.workspace != nullptr);
  if (rc != YP_OK) return rc;
  if (P.ws_bytes > d.workspace_bytes) {   // workspace too small for the split the planner wants: run unsplit
    if ((rc = plan_conv(d, &P, false)) != YP_OK) return rc;
  }
  ConvMaps maps;
  if ((rc = prepare_conv(d, P, d.workspace, maps)) != YP_OK) return rc;
  const ConvArgs& a = P.a;
  const int out_fmt = P.out_fmt;
  const dim3 grid = P.grid;
  const size_t smem = P.smem;
  const int units = P.units;
  const bool tf32 = P.tf32;
  if (a.persist == 2) return out_fmt == YP_FMT_F32X2 ? launch_drain<YP_FMT_F32X2>(maps, a, grid, smem, st) : launch_drain<YP_FMT_F32>(maps, a, grid, smem, st);
#define YP_DISPATCH_P(FMT, U) return launch_persist<FMT, U, false>(maps, a, grid, smem, st)
  if (a.persist) {
    if (out_fmt == YP_FMT_BF16) { if (units == 4) YP_DISPATCH_P(YP_FMT_BF16, 4); if (units == 2) YP_DISPATCH_P(YP_FMT_BF16, 2); if (units == 1) YP_DISPATCH_P(YP_FMT_BF16, 1); }
    else { if (units == 2) YP_DISPATCH_P(YP_FMT_F32, 2); if (units == 1) YP_DISPATCH_P(YP_FMT_F32, 1); }
  }
#undef YP_DISPATCH_P
#define YP_DISPATCH(FMT, U, TF) return P.nt == 512 ? launch<FMT, U, TF, 512>(maps, a, grid, smem, st) : launch<FMT, U, TF, 256>(maps, a, grid, smem, st)
  if (tf32) {
    if (out_fmt == YP_FMT_F32X2) { if (units == 2) YP_DISPATCH(YP_FMT_F32X2, 2, true); if (units == 1) YP_DISPATCH(YP_FMT_F32X2, 1, true); }
    else { if (units == 2) YP_DISPATCH(YP_FMT_F32, 2, true); if (units == 1) YP_DISPATCH(YP_FMT_F32, 1, true); }
  } else {
    if (out_fmt == YP_FMT_BF16) { if (units == 4) YP_DISPATCH(YP_FMT_BF16, 4, false); if (units == 2) YP_DISPATCH(YP_FMT_BF16, 2, false); if (units == 1) YP_DISPATCH(YP_FMT_BF16, 1, false); }
    else { if (units == 2) YP_DISPATCH(YP_FMT_F32, 2, false); if (units == 1) YP_DISPATCH(YP_FMT_F32, 1, false); }
  }
#undef YP_DISPATCH
  set_error("conv: no kernel instantiation for out_fmt=%d units=%d", out_fmt, units);
  return YP_ERR_SHAPE;
}

// ---------------------------------------------------------------------------------------------
// Layer chains (host side)
// ---------------------------------------------------------------------------------------------
namespace {

struct ChainKernel {
  ChainParams params;
  ConvMaps* d_maps = nullptr;      // [n_ops]
  unsigned* d_done = nullptr;      // [kMaxChain + 1] completion counters + exit counter
};
struct Chain {
  std::vector<ChainKernel*> kernels;
  std::vector<void*> allocs;
  size_t smem = 0;
  int n_items = 0, grid = 0, device = 0;
  ~Chain() {
    for (void* p : allocs) cudaFree(p);
    for (ChainKernel* k : kernels) delete k;
  }
};

int chain_variant(int out_fmt, int units) {
  if (out_fmt == YP_FMT_F32X2) return units == 2 ? 0 : (units == 1 ? 1 : -1);
  if (out_fmt == YP_FMT_F32) return units == 2 ? 2 : (units == 1 ? 3 : -1);
  return -1;
}

int chain_build(const YpChainOp* ops, int n_ops, Chain* ch) {
  YP_REQUIRE(get_encode() != nullptr, YP_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  YP_CUDA_OK(cudaGetDevice(&ch->device));
  static const int budget_kb = getenv("YP_CHAIN_SMEM_KB") ? atoi(getenv("YP_CHAIN_SMEM_KB")) : 160;
  const int G = sm_count();
  ch->grid = G;
  // ---- plan every operation (chain mode: one CTA per SM, common barrier offset)
  std::vector<ConvPlan> plans(n_ops);
  int bar_off = 0, max_nt = 16;
  size_t ws_total = 0, pool_smem = 0;
  for (int i = 0; i < n_ops; ++i) {
    const YpChainOp& op = ops[i];
    YP_REQUIRE(op.n_deps >= 0 && op.n_deps <= 8, YP_ERR_ARG, "chain: op %d has %d dependencies", i, op.n_deps);
    for (int k = 0; k < op.n_deps; ++k) YP_REQUIRE(op.deps[k] >= 0 && op.deps[k] < i, YP_ERR_ARG, "chain: op %d depends on op %d (not earlier in the list)", i, op.deps[k]);
    if (op.type == 1) {
      const YpView& v = op.pool;
      YP_REQUIRE(v.base && v.C % 4 == 0 && (v.C / 4) % SPPF_G == 0 && v.H * v.W <= 65535, YP_ERR_SHAPE, "chain: op %d: SPPF view unsupported", i);
      pool_smem = std::max(pool_smem, sppf_smem_bytes(v.H * v.W));
      continue;
    }
    YP_REQUIRE(op.type == 0, YP_ERR_ARG, "chain: op %d has unknown type %d", i, op.type);
    const YpConvDesc& d = op.conv;
    YP_REQUIRE(d.in.base && d.weight && d.n_out >= 1 && d.out[0].base, YP_ERR_ARG, "chain: op %d: null pointer in descriptor", i);
    YP_REQUIRE(d.algo == YP_ALGO_TCGEN05 && d.in.format == YP_FMT_F32X2 && !(d.epilogue & YP_EPI_ROWMIN), YP_ERR_SHAPE,
               "chain: op %d: only tcgen05 convolutions with F32X2 operands can be chained", i);
    g_chain_budget = budget_kb * 1024;
    const int rc = plan_conv(d, &plans[i], true);
    g_chain_budget = 0;
    if (rc != YP_OK) return rc;
    ConvPlan& P = plans[i];
    YP_REQUIRE(chain_variant(P.out_fmt, P.units) >= 0, YP_ERR_SHAPE, "chain: op %d: no epilogue instantiation for out_fmt=%d units=%d", i, P.out_fmt, P.units);
    bar_off = std::max(bar_off, P.a.bar_off);
    max_nt = std::max(max_nt, P.a.Nt);
    ws_total += (P.ws_bytes + 255) & ~static_cast<size_t>(255);
  }
  ch->smem = 1024 + bar_off + 8 * (2 * kMaxStages + 10) + max_nt * sizeof(float) + 16;
  ch->smem = std::max(ch->smem, pool_smem + 16);
  YP_REQUIRE(ch->smem <= 225 * 1024, YP_ERR_SHAPE, "chain: needs %zu bytes of shared memory", ch->smem);
  char* ws = nullptr;
  if (ws_total) {
    YP_CUDA_OK(cudaMalloc(&ws, ws_total));
    ch->allocs.push_back(ws);
    YP_CUDA_OK(cudaMemset(ws, 0, ws_total));
  }
  // ---- kernels of <= kMaxChain operations
  int cta_off = 0;
  for (int k0 = 0; k0 < n_ops; k0 += kMaxChain) {
    const int n = std::min(kMaxChain, n_ops - k0);
    ChainKernel* K = new ChainKernel();
    ch->kernels.push_back(K);
    memset(&K->params, 0, sizeof(K->params));
    K->params.n_ops = n;
    if (getenv("YP_CHAIN_DEBUG")) {
      const size_t bytes = static_cast<size_t>(n) * G * 16 * sizeof(long long);
      YP_CUDA_OK(cudaMalloc(&K->params.dbg, bytes));
      ch->allocs.push_back(K->params.dbg);
      YP_CUDA_OK(cudaMemset(K->params.dbg, 0, bytes));
    }
    std::vector<ConvMaps> maps(n);
    int n_pool = 0;
    for (int j = 0; j < n; ++j) {
      const int i = k0 + j;
      const YpChainOp& op = ops[i];
      ChainMeta& m = K->params.meta[j];
      m.type = op.type;
      if (op.type == 1) {
        YP_REQUIRE(n_pool < 2, YP_ERR_SHAPE, "chain: more than two SPPF poolings in one kernel");
        ChainPool& pp = K->params.pool[n_pool];
        pp.cat4 = op.pool; pp.C = op.pool.C / 4; pp.groups = pp.C / SPPF_G;
        m.variant = n_pool++;
        m.gx = pp.groups; m.gy = op.pool.B;
        m.n_items = pp.groups * op.pool.B;
        memset(&maps[j], 0, sizeof(ConvMaps));
      } else {
        ConvPlan& P = plans[i];
        P.a.bar_off = bar_off;
        void* w = nullptr;
        if (P.ws_bytes) { w = ws; ws += (P.ws_bytes + 255) & ~static_cast<size_t>(255); }
        const int rc = prepare_conv(op.conv, P, w, maps[j]);
        if (rc != YP_OK) return rc;
        P.a.dbg = nullptr; P.a.ablate = 0;
        K->params.args[j] = P.a;
        m.variant = chain_variant(P.out_fmt, P.units);
        m.gx = P.grid.x; m.gy = P.grid.y;
        m.n_items = P.grid.x * P.grid.y * P.grid.z;
      }
      m.cta_off = cta_off;
      cta_off = (cta_off + m.n_items) % G;
      ch->n_items += m.n_items;
      m.n_deps = 0;
      for (int k = 0; k < op.n_deps; ++k) {
        const int dj = op.deps[k] - k0;
        if (dj < 0) continue;                       // produced by an earlier kernel of the chain: ordered by the stream
        YP_REQUIRE(m.n_deps < kMaxChainDeps, YP_ERR_SHAPE, "chain: op %d has more than %d dependencies inside one kernel", i, kMaxChainDeps);
        m.dep[m.n_deps] = dj;
        m.dep_target[m.n_deps] = K->params.meta[dj].n_items;
        ++m.n_deps;
      }
    }
    YP_CUDA_OK(cudaMalloc(&K->d_maps, n * sizeof(ConvMaps)));
    ch->allocs.push_back(K->d_maps);
    YP_CUDA_OK(cudaMemcpy(K->d_maps, maps.data(), n * sizeof(ConvMaps), cudaMemcpyHostToDevice));
    YP_CUDA_OK(cudaMalloc(&K->d_done, (kMaxChain + 1) * sizeof(unsigned)));
    ch->allocs.push_back(K->d_done);
    YP_CUDA_OK(cudaMemset(K->d_done, 0, (kMaxChain + 1) * sizeof(unsigned)));
  }
  YP_CUDA_OK(cudaDeviceSynchronize());
  return YP_OK;
}

}  // namespace

int conv_chain_create(const YpChainOp* ops, int n_ops, void** out) {
  Chain* ch = new Chain();
  const int rc = chain_build(ops, n_ops, ch);
  if (rc != YP_OK) { delete ch; return rc; }
  *out = ch;
  return YP_OK;
}

int conv_chain_launch(void* chain, cudaStream_t st) {
  Chain* ch = static_cast<Chain*>(chain);
  auto kern = conv_chain_kernel<true, kChainThreads>;
  static thread_local size_t configured = 0;
  if (ch->smem > configured) {   // the kernel also has ~1 KB of static shared memory: ask for what the chain needs, not for the maximum
    YP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ch->smem)));
    configured = ch->smem;
  }
  static const bool coop = getenv("YP_CHAIN_NO_COOP") == nullptr;
  for (ChainKernel* K : ch->kernels) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ch->grid); cfg.blockDim = dim3(kChainThreads); cfg.dynamicSmemBytes = ch->smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the in-kernel waits cannot starve an unscheduled CTA
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = coop ? 1 : 0;
    YP_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, K->params, static_cast<const ConvMaps*>(K->d_maps), K->d_done));
  }
  return YP_OK;
}

// debug: device pointer of kernel k's timeline ([n_ops][#CTAs][16] int64 nanosecond stamps: 0 item start, 6 dependencies met, 1 prologue
// done, 2 accumulators complete, 3 stores complete, 4 item end, 7 completion published), its op count and the CTA count
int conv_chain_debug(void* chain, int k, void** dbg, int* n_ops, int* n_ctas) {
  Chain* ch = static_cast<Chain*>(chain);
  YP_REQUIRE(k >= 0 && k < static_cast<int>(ch->kernels.size()), YP_ERR_ARG, "chain: kernel index %d out of range", k);
  *dbg = ch->kernels[k]->params.dbg; *n_ops = ch->kernels[k]->params.n_ops; *n_ctas = ch->grid;
  return YP_OK;
}

int conv_chain_destroy(void* chain) {
  delete static_cast<Chain*>(chain);
  return YP_OK;
}

int conv_chain_info(void* chain, int* n_kernels, int* n_items, int* smem) {
  Chain* ch = static_cast<Chain*>(chain);
  if (n_kernels) *n_kernels = static_cast<int>(ch->kernels.size());
  if (n_items) *n_items = ch->n_items;
  if (smem) *smem = static_cast<int>(ch->smem);
  return YP_OK;
}

}  // namespace yp
