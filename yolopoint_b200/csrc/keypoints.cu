// Keypoint heatmap (65-way cell softmax + depth-to-space) and exact parallel greedy keypoint NMS
// (utils/utils.py:118-182, 232-262, 465-485 and demo.py:138-198 of the reference).
#include "common.cuh"

namespace yp {
namespace {

// ------------------------------------------------------------------------------------------------
// heatmap: one thread per 8x8 cell, 65 logits in registers
// ------------------------------------------------------------------------------------------------
// The kernel is bound by instruction issue, not by memory (measured at batch 32, 1280x736: scalar vs float4 loads and whole-sector
// vs half-sector stores made no difference, the 64 IEEE divisions per cell did): the normalisation is one reciprocal per cell and
// a multiply per pixel (<= 1 ulp from the quotient; the softmax itself is only reproducible to ~1e-7 across exp implementations).
// kVec: channels contiguous and 16-byte aligned per cell (the engine's NHWC `semi` buffer with 80-channel rows) -> 16 float4 + 1
// scalar load instead of 65 scalar loads that each touch 32 different sectors.
template <bool kVec>
__global__ void __launch_bounds__(128) heatmap_kernel(const float* __restrict__ semi, int B, int Hc, int Wc, long long sB, long long sC,
                                                       long long sH, long long sW, int variant, float* __restrict__ heat) {
  const int64_t total = static_cast<int64_t>(B) * Hc * Wc;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int wc = static_cast<int>(idx % Wc);
  const int hc = static_cast<int>((idx / Wc) % Hc);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(Wc) * Hc));
  const float* s = semi + b * sB + hc * sH + wc * sW;
  float v[65];
  float mx = -INFINITY;
  if (kVec) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 t = __ldg(s4 + c);
      v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
    }
    v[64] = __ldg(s + 64);
#pragma unroll
    for (int c = 0; c < 65; ++c) mx = fmaxf(mx, v[c]);
  } else {
#pragma unroll
    for (int c = 0; c < 65; ++c) { v[c] = s[c * sC]; mx = fmaxf(mx, v[c]); }
  }
  float sum = 0.0f;
  if (variant == 0) {  // torch.softmax: exp(x - max) / sum
#pragma unroll
    for (int c = 0; c < 65; ++c) { v[c] = expf(__fsub_rn(v[c], mx)); sum = __fadd_rn(sum, v[c]); }
  } else {  // demo.py:140-141: exp(x) / (sum + 1e-5)
#pragma unroll
    for (int c = 0; c < 65; ++c) { v[c] = expf(v[c]); sum = __fadd_rn(sum, v[c]); }
    sum = __fadd_rn(sum, 0.00001f);
  }
  const float inv = __fdiv_rn(1.0f, sum);
  const int W = Wc * 8;
  float* o = heat + (static_cast<int64_t>(b) * Hc * 8 + hc * 8) * W + wc * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 lo4 = make_float4(__fmul_rn(v[8 * i], inv), __fmul_rn(v[8 * i + 1], inv), __fmul_rn(v[8 * i + 2], inv), __fmul_rn(v[8 * i + 3], inv));
    const float4 hi4 = make_float4(__fmul_rn(v[8 * i + 4], inv), __fmul_rn(v[8 * i + 5], inv), __fmul_rn(v[8 * i + 6], inv), __fmul_rn(v[8 * i + 7], inv));
    reinterpret_cast<float4*>(o + static_cast<int64_t>(i) * W)[0] = lo4;
    reinterpret_cast<float4*>(o + static_cast<int64_t>(i) * W)[1] = hi4;
  }
}

// ------------------------------------------------------------------------------------------------
// keypoint NMS.  state per pixel: 0 = not a candidate, 1 = undecided, 2 = kept, 3 = suppressed.
// Sequential reference: visit candidates by (confidence desc, raster index asc); a candidate is kept iff
// no already-kept candidate lies within Chebyshev distance r.  Parallel fixed point with the same result:
//   p becomes SUPPRESSED as soon as a higher-priority candidate in its window is KEPT,
//   p becomes KEPT as soon as every higher-priority candidate in its window is SUPPRESSED.
// Every decision is final and equals the sequential one (induction over priority), so in-place
// asynchronous updates are safe; rounds repeat (grid-wide barrier) until nothing is undecided.
// ------------------------------------------------------------------------------------------------
struct KpWs {
  unsigned char* state;        // [B*H*W]
  unsigned int* tile_undecided;  // [B * tiles]
  unsigned int* round_total;     // [KP_ROUNDS] undecided candidates left after each round
  int* n_list;                 // [B]
  int* n_thresh;               // [B]  pixels with heat >= conf_thresh (candidates before the NMS), right behind n_list
  unsigned long long* list;    // [B][max_pts]  (ordered_conf << 32) | raster index
};


__device__ __forceinline__ unsigned int ordered_bits(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unordered_bits(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// Tile-local formulation.  Each CTA owns a 32x32-pixel tile; per round it loads the tile plus an r-wide halo of states
// into shared memory, sorts the tile's undecided candidates by priority (rank sort) and lets one warp resolve them IN
// PRIORITY ORDER -- inside a tile that is the reference's sequential scan, so long dependency chains cost shared-memory
// latency.  Only decisions that depend on a still-undecided halo candidate of a neighbouring tile are deferred to the
// next round.  Rounds are separate (ordinary, non-cooperative) launches, so any number of these pipelines can run
// concurrently on one GPU; tiles without undecided candidates exit at once; a final single-CTA kernel sweeps until
// nothing is undecided, which bounds the launch count without giving up exactness.
constexpr int KT = 32;       // tile edge
constexpr int KR_MAX = 16;   // largest supported nms_dist
constexpr int KP_ROUNDS = 8; // parallel rounds before the sequential sweep (dense noise needs ~7)

constexpr int kWinSteps = ((2 * KR_MAX + 1) * (2 * KR_MAX + 1) + 31) / 32;

struct KpTileSmem {
  float* sheat; unsigned long long* keys; unsigned short* cand; unsigned short* sorted; unsigned char* sstate; int* n_list;
  const short* win_off;   // [kWinSteps * 32] cell offsets of the (2r+1)^2 window positions, see kp_window_offsets
};

// window position idx (row-major over the (2r+1)^2 window) -> offset in the (KT + 2r)-wide shared-memory tile; 0 for the centre
// and for the padding positions >= (2r+1)^2.  Computed once per kernel by the whole CTA.
__device__ __forceinline__ void kp_window_offsets(short* tab, int r) {
  const int win = 2 * r + 1, RW = KT + 2 * r;
  for (int idx = threadIdx.x; idx < kWinSteps * 32; idx += blockDim.x)
    tab[idx] = idx < win * win ? static_cast<short>((idx / win - r) * RW + (idx % win - r)) : static_cast<short>(0);
}

// One round for tile t.  Returns (to thread 0 only... every thread gets its partial) the number of still-undecided
// interior candidates counted by this thread.
__device__ unsigned int kp_process_tile(const float* __restrict__ heat, int B, int H, int W, float thr, int r, const KpWs& ws, int t,
                                        bool first, const KpTileSmem& sm) {
  const int RW = KT + 2 * r, RN = RW * RW;
  const int tiles_x = (W + KT - 1) / KT, tiles_y = (H + KT - 1) / KT;
  const int64_t HW = static_cast<int64_t>(H) * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int win = 2 * r + 1;
  const int b = t / (tiles_x * tiles_y);
  const int tr = t - b * tiles_x * tiles_y;
  const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
  const int y0 = ty * KT - r, x0 = tx * KT - r;
  const float* hb = heat + b * HW;
  unsigned char* sb = ws.state + b * HW;
  for (int c = threadIdx.x; c < RN; c += blockDim.x) {
    const int ly = c / RW, lx = c - ly * RW;
    const int gy = y0 + ly, gx = x0 + lx;
    const bool inside = gy >= 0 && gy < H && gx >= 0 && gx < W;
    const float h = inside ? hb[gy * W + gx] : -INFINITY;
    sm.sheat[c] = h;
    unsigned char st = 0;
    if (inside) st = first ? (h >= thr ? 1 : 0) : __ldcg(sb + gy * W + gx);
    sm.sstate[c] = st;
  }
  if (threadIdx.x == 0) *sm.n_list = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < KT * KT; c += blockDim.x) {
    const int ly = c / KT + r, lx = c % KT + r;
    const int cell = ly * RW + lx;
    if (sm.sstate[cell] == 1) {
      const int slot = atomicAdd(sm.n_list, 1);
      const unsigned int raster = static_cast<unsigned int>((y0 + ly) * W + (x0 + lx));
      sm.keys[slot] = (static_cast<unsigned long long>(ordered_bits(sm.sheat[cell])) << 32) | (0xffffffffu - raster);
      sm.cand[slot] = static_cast<unsigned short>(cell);
    }
  }
  __syncthreads();
  const int n = *sm.n_list;
  if (first && threadIdx.x == 0 && n) atomicAdd(&ws.n_thresh[b], n);   // round 0 sees every pixel >= threshold of its tile exactly once
  for (int i = threadIdx.x; i < n; i += blockDim.x) {  // rank sort, descending priority (keys are unique)
    const unsigned long long ki = sm.keys[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += sm.keys[j] > ki ? 1 : 0;
    sm.sorted[rank] = sm.cand[i];
  }
  __syncthreads();
  {
    // Resolve the tile's undecided candidates: warp w walks candidates w, w + nwarps, ... of the priority-sorted list, so the whole
    // CTA sweeps the list roughly in priority order; sweeps repeat until one of them decides nothing (tile-local fixed point; what
    // is left depends on an undecided halo candidate of a neighbouring tile and waits for the next round).  Each decision is the
    // sequential algorithm's decision whatever the interleaving: p is SUPPRESSED iff a kept candidate lies in its window, KEPT iff
    // every higher-priority candidate in its window is decided and none is kept -- both are final, so a stale read of a
    // neighbour's state can only postpone a decision to the next sweep.  A kept candidate marks the undecided cells of its window
    // suppressed (nms_fast's own step, utils/utils.py:165-170; they all have lower priority, or p could not have been kept), so
    // most candidates cost one shared-memory read.  The (2r+1)^2 window is walked in 32-cell steps through a table of cell
    // offsets built once per kernel.  (The first version resolved with ONE warp and divided by the window width per cell: 60 us
    // per round; ncu showed the other seven warps waiting at the barrier for 73 % of the kernel.)
    // Priority between two cells of one tile: confidence, then the smaller cell index (= the smaller raster index).
    const int steps = (win * win + 31) / 32;
    const int nwarps = blockDim.x >> 5;
    const short* off = sm.win_off + lane;      // off[32 t]: cell offset of window position lane + 32 t (0 = the candidate itself / padding)
    volatile unsigned char* vstate = sm.sstate;
    for (;;) {
      __syncthreads();
      if (threadIdx.x == 0) *sm.n_list = 0;    // reused as "this sweep decided something"
      __syncthreads();
      bool progress = false;
      for (int i = warp; i < n; i += nwarps) {
        const int p = sm.sorted[i];
        if (vstate[p] != 1) continue;           // decided (uniform: same address)
        const float hp = sm.sheat[p];
        bool kept = false, pend = false;
        for (int t = 0; t < steps; ++t) {
          const int o = off[32 * t];
          if (o != 0) {
            const int q = p + o;
            const unsigned char sq = vstate[q];
            if (sq == 2) kept = true;
            else if (sq == 1) {
              const float hq = sm.sheat[q];
              if (hq > hp || (hq == hp && q < p)) pend = true;
            }
          }
        }
        kept = __any_sync(0xffffffffu, kept);
        pend = __any_sync(0xffffffffu, pend);
        if (kept) {
          if (lane == 0) vstate[p] = 3;
          progress = true;
        } else if (!pend) {                      // keep p: everything still undecided in its window loses to it
          if (lane == 0) vstate[p] = 2;
          for (int t = 0; t < steps; ++t) {
            const int o = off[32 * t];
            if (o != 0 && vstate[p + o] == 1) vstate[p + o] = 3;
          }
          progress = true;
        }
        __syncwarp();
      }
      if (progress && lane == 0) *sm.n_list = 1;
      __syncthreads();
      if (*sm.n_list == 0) break;
    }
  }
  __syncthreads();
  unsigned int undecided = 0;
  if (first) {   // publish the whole interior: neighbours read the candidates' states from global memory in later rounds
    for (int c = threadIdx.x; c < KT * KT; c += blockDim.x) {
      const int ly = c / KT + r, lx = c % KT + r;
      const int gy = y0 + ly, gx = x0 + lx;
      if (gy < H && gx < W) {
        const unsigned char st = sm.sstate[ly * RW + lx];
        sb[gy * W + gx] = st;
        undecided += st == 1 ? 1u : 0u;
      }
    }
  } else {       // later rounds: only the candidates that were undecided can have changed
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int p = sm.sorted[i];
      const int py = p / RW, px = p - py * RW;
      const unsigned char st = sm.sstate[p];
      if (st != 1) sb[(y0 + py) * W + (x0 + px)] = st; else ++undecided;
    }
  }
  __syncthreads();
  return undecided;
}

__device__ __forceinline__ KpTileSmem kp_carve(unsigned char* base, int r, int* n_list, const short* win_off) {
  const int RW = KT + 2 * r, RN = RW * RW;
  KpTileSmem sm;
  sm.sheat = reinterpret_cast<float*>(base);
  sm.keys = reinterpret_cast<unsigned long long*>(base + ((RN * 4 + 7) & ~7));
  sm.cand = reinterpret_cast<unsigned short*>(sm.keys + KT * KT);
  sm.sorted = sm.cand + KT * KT;
  sm.sstate = reinterpret_cast<unsigned char*>(sm.sorted + KT * KT);
  sm.n_list = n_list;
  sm.win_off = win_off;
  return sm;
}

// ws.tile_undecided[t]: undecided interior candidates of tile t after its last processed round
__global__ void __launch_bounds__(256) kp_round_kernel(const float* __restrict__ heat, int B, int H, int W, float thr, int r, KpWs ws, int first, int round) {
  extern __shared__ unsigned char kp_smem[];
  __shared__ int n_list;
  __shared__ unsigned int total;
  __shared__ short win_off[kWinSteps * 32];
  const int t = blockIdx.x;
  if (!first && ws.tile_undecided[t] == 0) return;
  kp_window_offsets(win_off, r);
  const KpTileSmem sm = kp_carve(kp_smem, r, &n_list, win_off);
  if (threadIdx.x == 0) total = 0;
  unsigned int u = kp_process_tile(heat, B, H, W, thr, r, ws, t, first != 0, sm);
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) u += __shfl_xor_sync(0xffffffffu, u, sft);
  if ((threadIdx.x & 31) == 0 && u) atomicAdd(&total, u);
  __syncthreads();
  if (threadIdx.x == 0) {
    ws.tile_undecided[t] = total;
    if (total) atomicAdd(&ws.round_total[round], total);
  }
}

// sequential safety net: one CTA sweeps the tiles that still have undecided candidates until none is left
__global__ void __launch_bounds__(256) kp_sweep_kernel(const float* __restrict__ heat, int B, int H, int W, float thr, int r, KpWs ws, int T) {
  extern __shared__ unsigned char kp_smem[];
  __shared__ int n_list;
  __shared__ unsigned int total, any;
  __shared__ short win_off[kWinSteps * 32];
  if (ws.round_total[KP_ROUNDS - 1] == 0) return;   // the parallel rounds finished everything (the usual case)
  kp_window_offsets(win_off, r);
  const KpTileSmem sm = kp_carve(kp_smem, r, &n_list, win_off);
  for (;;) {
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    for (int t = 0; t < T; ++t) {
      if (ws.tile_undecided[t] == 0) continue;   // uniform: written by thread 0 before a barrier, read after it
      if (threadIdx.x == 0) total = 0;
      unsigned int u = kp_process_tile(heat, B, H, W, thr, r, ws, t, false, sm);
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) u += __shfl_xor_sync(0xffffffffu, u, sft);
      if ((threadIdx.x & 31) == 0 && u) atomicAdd(&total, u);
      __syncthreads();
      if (threadIdx.x == 0) { ws.tile_undecided[t] = total; any |= total; }
      __threadfence();
      __syncthreads();
    }
    if (any == 0) break;
    __syncthreads();
  }
}

// python slice semantics of mask[y0:y1, x0:x1] on an axis of length L (demo.py:185-186)
__device__ __forceinline__ void py_slice(int a, int b, int L, int* s, int* e) {
  *s = a < 0 ? max(a + L, 0) : min(a, L);
  *e = b < 0 ? max(b + L, 0) : min(b, L);
}

// survivors -> border filter -> box filter -> unordered list of packed keys
__global__ void kp_collect_kernel(const float* __restrict__ heat, int B, int H, int W, int border, const float* __restrict__ boxes,
                                  const int* __restrict__ box_count, int box_ld, int max_pts, KpWs ws) {
  extern __shared__ int sbox[];  // [nb][4] slice bounds x0,x1,y0,y1
  const int b = blockIdx.y;
  int nb = 0;
  if (boxes) {
    nb = box_count[b];
    if (nb < 0) nb = 0;  // overflowed NMS: no boxes are trusted
    if (nb > box_ld) nb = box_ld;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
      const float* bx = boxes + (static_cast<int64_t>(b) * box_ld + i) * 6;
      int sx, ex, sy, ey;
      py_slice(static_cast<int>(rintf(bx[0])), static_cast<int>(rintf(bx[2])), W, &sx, &ex);
      py_slice(static_cast<int>(rintf(bx[1])), static_cast<int>(rintf(bx[3])), H, &sy, &ey);
      sbox[4 * i] = sx; sbox[4 * i + 1] = ex; sbox[4 * i + 2] = sy; sbox[4 * i + 3] = ey;
    }
  }
  __syncthreads();
  const int HW = H * W;
  for (int pp = blockIdx.x * blockDim.x + threadIdx.x; pp < HW; pp += gridDim.x * blockDim.x) {
    if (ws.state[static_cast<int64_t>(b) * HW + pp] != 2) continue;
    const int y = pp / W, x = pp - y * W;
    if (x < border || x >= W - border || y < border || y >= H - border) continue;
    bool inside = false;
    for (int i = 0; i < nb && !inside; ++i)
      inside = x >= sbox[4 * i] && x < sbox[4 * i + 1] && y >= sbox[4 * i + 2] && y < sbox[4 * i + 3];
    if (inside) continue;
    const int slot = atomicAdd(&ws.n_list[b], 1);
    if (slot < max_pts)
      ws.list[static_cast<int64_t>(b) * max_pts + slot] =
          (static_cast<unsigned long long>(ordered_bits(heat[static_cast<int64_t>(b) * HW + pp])) << 32) | static_cast<unsigned int>(pp);
  }
}

// rank sort by packed key descending (confidence desc, ties: larger raster index first) and emit (x, y, conf)
__global__ void kp_emit_kernel(int W, int max_pts, KpWs ws, float* __restrict__ out_pts, int* __restrict__ out_count) {
  __shared__ unsigned long long tile[256];
  const int b = blockIdx.y;
  const int nall = ws.n_list[b];
  const int n = min(nall, max_pts);
  if (blockIdx.x == 0 && threadIdx.x == 0) out_count[b] = nall > max_pts ? -1 - nall : nall;
  if (blockIdx.x * blockDim.x >= n) return;
  const unsigned long long* list = ws.list + static_cast<int64_t>(b) * max_pts;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long ki = i < n ? list[i] : 0ull;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    const int j = j0 + threadIdx.x;
    tile[threadIdx.x] = j < n ? list[j] : 0ull;
    __syncthreads();
    const int lim = min(256, n - j0);
    for (int k = 0; k < lim; ++k) rank += tile[k] > ki ? 1 : 0;  // keys are unique (distinct raster index)
    __syncthreads();
  }
  if (i < n) {
    const unsigned int pp = static_cast<unsigned int>(ki & 0xffffffffull);
    float* o = out_pts + (static_cast<int64_t>(b) * max_pts + rank) * 3;
    o[0] = static_cast<float>(pp % W);
    o[1] = static_cast<float>(pp / W);
    o[2] = unordered_bits(static_cast<unsigned int>(ki >> 32));
  }
}

// ------------------------------------------------------------------------------------------------
// Whole-frame pipeline, critical path after the box NMS: the keypoint list (NMS survivors inside the border, confidence
// descending) was built while the detection branch was still running; what is left is the in-box filter (demo.py:178-198) as an
// ORDER-PRESERVING compaction.  One CTA per image: box slices in shared memory, a block-wide exclusive scan per 1024 points.
// Also initialises the keys of the two-way match that follows (previous frame x this frame).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) kp_filter_kernel(const float* __restrict__ pts_all, const int* __restrict__ n_all, int max_pts, int H, int W,
                                                         const float* __restrict__ boxes, const int* __restrict__ box_count, int box_ld,
                                                         float* __restrict__ out_pts, int* __restrict__ out_sel, int* __restrict__ out_count,
                                                         const int* __restrict__ n_prev, unsigned long long* __restrict__ row_key,
                                                         unsigned long long* __restrict__ col_key) {
  extern __shared__ int4 sbox4[];  // [nb] slice bounds x0,x1,y0,y1 (padded to a multiple of 4 with empty slices)
  int* sbox = reinterpret_cast<int*>(sbox4);
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int nb = 0;
  if (boxes) {
    nb = min(max(box_count[b], 0), box_ld);
    for (int i = tid; i < nb; i += blockDim.x) {
      const float* bx = boxes + (static_cast<int64_t>(b) * box_ld + i) * 6;
      int sx, ex, sy, ey;
      py_slice(static_cast<int>(rintf(bx[0])), static_cast<int>(rintf(bx[2])), W, &sx, &ex);
      py_slice(static_cast<int>(rintf(bx[1])), static_cast<int>(rintf(bx[3])), H, &sy, &ey);
      sbox[4 * i] = sx; sbox[4 * i + 1] = ex; sbox[4 * i + 2] = sy; sbox[4 * i + 3] = ey;
    }
    if (tid < 4 && nb + tid < ((nb + 3) & ~3)) sbox4[nb + tid] = make_int4(0, 0, 0, 0);   // empty slices: never contain a point
  }
  if (tid == 0) carry_s = 0;
  __syncthreads();
  const int n = min(max(n_all[b], 0), max_pts);
  const float* pin = pts_all + static_cast<int64_t>(b) * max_pts * 3;
  float* pout = out_pts + static_cast<int64_t>(b) * max_pts * 3;
  int* sel = out_sel + static_cast<int64_t>(b) * max_pts;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + tid;
    bool keep = false;
    float px = 0.f, py = 0.f, pc = 0.f;
    if (i < n) {
      px = pin[i * 3]; py = pin[i * 3 + 1]; pc = pin[i * 3 + 2];
      const int x = static_cast<int>(px), y = static_cast<int>(py);
      bool inside = false;   // four boxes per step, no early exit: independent 16-byte loads instead of a dependent branch chain
      for (int j = 0; j < nb; j += 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int4 s4 = sbox4[j + e];
          inside |= x >= s4.x && x < s4.y && y >= s4.z && y < s4.w;
        }
      }
      keep = !inside;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_sum[warp] = __popc(m);
    __syncthreads();
    const int carry = carry_s;
    int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_sum[w];
    if (keep) {
      const int pos = carry + before + __popc(m & ((1u << lane) - 1u));
      pout[pos * 3] = px; pout[pos * 3 + 1] = py; pout[pos * 3 + 2] = pc;
      sel[pos] = i;
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < (blockDim.x >> 5); ++w) tot += warp_sum[w];
      carry_s = carry + tot;
    }
    __syncthreads();
  }
  const int kept = carry_s;
  if (tid == 0) out_count[b] = n_all[b] < 0 ? n_all[b] : kept;   // an overflowed list stays flagged
  if (row_key) {
    const int np = n_prev ? min(max(n_prev[b], 0), max_pts) : 0;
    for (int i = tid; i < np; i += blockDim.x) row_key[static_cast<int64_t>(b) * max_pts + i] = ~0ull;
    for (int i = tid; i < kept; i += blockDim.x) col_key[static_cast<int64_t>(b) * max_pts + i] = ~0ull;
  }
}

size_t align_up(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

size_t carve(KpWs* ws, char* base, int B, int H, int W, int max_pts) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
  ws->state = reinterpret_cast<unsigned char*>(take(static_cast<size_t>(B) * H * W));
  ws->tile_undecided = reinterpret_cast<unsigned int*>(take(sizeof(unsigned int) * B * ((H + 31) / 32) * ((W + 31) / 32)));
  ws->round_total = reinterpret_cast<unsigned int*>(take(sizeof(unsigned int) * 16));
  ws->n_list = reinterpret_cast<int*>(take(sizeof(int) * 2 * B));
  ws->n_thresh = ws->n_list ? ws->n_list + B : nullptr;
  ws->list = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * B * max_pts));
  return off;
}

}  // namespace
}  // namespace yp

extern "C" int yp_heatmap(const float* semi, int32_t B, int32_t Hc, int32_t Wc, int64_t sB, int64_t sC, int64_t sH, int64_t sW,
                          int32_t variant, float* heat, void* stream) {
  YP_REQUIRE(semi && heat, YP_ERR_ARG, "heatmap: null pointer");
  YP_REQUIRE(B > 0 && Hc > 0 && Wc > 0 && (variant == 0 || variant == 1), YP_ERR_SHAPE, "heatmap: bad shape/variant");
  YP_REQUIRE(yp::aligned16(heat), YP_ERR_ALIGN, "heatmap: output not 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * Hc * Wc;
  const bool vec = sC == 1 && sB % 4 == 0 && sH % 4 == 0 && sW % 4 == 0 && yp::aligned16(semi);   // NHWC rows, 16-byte aligned cells
  const unsigned blocks = static_cast<unsigned>(yp::ceil_div64(total, 128));
  if (vec) yp::heatmap_kernel<true><<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(semi, B, Hc, Wc, sB, sC, sH, sW, variant, heat);
  else yp::heatmap_kernel<false><<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(semi, B, Hc, Wc, sB, sC, sH, sW, variant, heat);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" size_t yp_keypoints_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t max_pts) {
  if (B <= 0 || H <= 0 || W <= 0 || max_pts <= 0) return 0;
  yp::KpWs ws;
  return yp::carve(&ws, nullptr, B, H, W, max_pts);
}

extern "C" int yp_keypoints_nms(const float* heat, int32_t B, int32_t H, int32_t W, float conf_thresh, int32_t nms_dist, int32_t max_pts,
                                void* workspace, size_t workspace_bytes, void* stream) {
  YP_REQUIRE(heat && workspace, YP_ERR_ARG, "keypoints: null pointer");
  YP_REQUIRE(B > 0 && H > 0 && W > 0 && max_pts > 0 && nms_dist >= 0, YP_ERR_SHAPE, "keypoints: bad shape");
  YP_REQUIRE(static_cast<int64_t>(H) * W < (1ll << 31), YP_ERR_SHAPE, "keypoints: image too large");
  YP_REQUIRE(nms_dist <= yp::KR_MAX, YP_ERR_SHAPE, "keypoints: nms_dist=%d exceeds the supported maximum %d", nms_dist, yp::KR_MAX);
  yp::KpWs ws;
  const size_t need = yp::carve(&ws, static_cast<char*>(workspace), B, H, W, max_pts);
  YP_REQUIRE(workspace_bytes >= need, YP_ERR_CAPACITY, "keypoints: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int RW = yp::KT + 2 * nms_dist;
  const size_t nms_smem = ((static_cast<size_t>(RW) * RW * 4 + 7) & ~static_cast<size_t>(7)) + yp::KT * yp::KT * (8 + 2 + 2) + static_cast<size_t>(RW) * RW;
  static thread_local bool raised = false;
  if (nms_smem > 48 * 1024 && !raised) {
    YP_CUDA_OK(cudaFuncSetAttribute(yp::kp_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    YP_CUDA_OK(cudaFuncSetAttribute(yp::kp_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    raised = true;
  }
  const int tiles = B * yp::ceil_div(H, yp::KT) * yp::ceil_div(W, yp::KT);
  YP_CUDA_OK(cudaMemsetAsync(ws.round_total, 0, yp::align_up(sizeof(unsigned int) * 16) + sizeof(int) * 2 * B, st));   // round_total, n_list, n_thresh (adjacent)
  for (int round = 0; round < yp::KP_ROUNDS; ++round)
    yp::kp_round_kernel<<<tiles, 256, nms_smem, st>>>(heat, B, H, W, conf_thresh, nms_dist, ws, round == 0 ? 1 : 0, round);
  yp::kp_sweep_kernel<<<1, 256, nms_smem, st>>>(heat, B, H, W, conf_thresh, nms_dist, ws, tiles);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_keypoints_collect(const float* heat, int32_t B, int32_t H, int32_t W, int32_t border, const float* boxes,
                                    const int32_t* box_count, int32_t box_ld, float* out_pts, int32_t* out_count, int32_t max_pts,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  YP_REQUIRE(heat && out_pts && out_count && workspace, YP_ERR_ARG, "keypoints: null pointer");
  YP_REQUIRE(B > 0 && H > 0 && W > 0 && max_pts > 0 && border >= 0, YP_ERR_SHAPE, "keypoints: bad shape");
  YP_REQUIRE(!boxes || (box_count && box_ld > 0 && box_ld <= 8192), YP_ERR_ARG, "keypoints: boxes need box_count and 0 < box_ld <= 8192");
  yp::KpWs ws;
  const size_t need = yp::carve(&ws, static_cast<char*>(workspace), B, H, W, max_pts);
  YP_REQUIRE(workspace_bytes >= need, YP_ERR_CAPACITY, "keypoints: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = boxes ? sizeof(int) * 4 * box_ld : 0;
  if (smem > 48 * 1024) {
    static thread_local bool raised = false;
    if (!raised) { YP_CUDA_OK(cudaFuncSetAttribute(yp::kp_collect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024)); raised = true; }
  }
  const int HW = H * W;
  int cblocks = yp::ceil_div(HW, 256);
  if (cblocks > 4 * yp::sm_count()) cblocks = 4 * yp::sm_count();
  yp::kp_collect_kernel<<<dim3(cblocks, B), 256, smem, st>>>(heat, B, H, W, border, boxes, box_count, box_ld, max_pts, ws);
  yp::kp_emit_kernel<<<dim3(yp::ceil_div(max_pts, 256), B), 256, 0, st>>>(W, max_pts, ws, out_pts, out_count);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_keypoints(const float* heat, int32_t B, int32_t H, int32_t W, float conf_thresh, int32_t nms_dist, int32_t border,
                            const float* boxes, const int32_t* box_count, int32_t box_ld, float* out_pts, int32_t* out_count,
                            int32_t max_pts, void* workspace, size_t workspace_bytes, void* stream) {
  const int rc = yp_keypoints_nms(heat, B, H, W, conf_thresh, nms_dist, max_pts, workspace, workspace_bytes, stream);
  if (rc != YP_OK) return rc;
  return yp_keypoints_collect(heat, B, H, W, border, boxes, box_count, box_ld, out_pts, out_count, max_pts, workspace, workspace_bytes, stream);
}

extern "C" int yp_keypoints_filter(const float* pts_all, const int32_t* n_all, int32_t B, int32_t max_pts, int32_t H, int32_t W,
                                   const float* boxes, const int32_t* box_count, int32_t box_ld, float* out_pts, int32_t* out_sel,
                                   int32_t* out_count, const int32_t* n_prev, unsigned long long* row_key, unsigned long long* col_key,
                                   void* stream) {
  YP_REQUIRE(pts_all && n_all && out_pts && out_sel && out_count, YP_ERR_ARG, "keypoints_filter: null pointer");
  YP_REQUIRE(B > 0 && max_pts > 0 && H > 0 && W > 0, YP_ERR_SHAPE, "keypoints_filter: bad shape");
  YP_REQUIRE(!boxes || (box_count && box_ld > 0 && box_ld <= 8192), YP_ERR_ARG, "keypoints_filter: boxes need box_count and 0 < box_ld <= 8192");
  YP_REQUIRE(!row_key || col_key, YP_ERR_ARG, "keypoints_filter: row_key needs col_key");
  const size_t smem = boxes ? sizeof(int) * 4 * (box_ld + 4) : 0;
  if (smem > 40 * 1024) {
    static thread_local bool raised = false;
    if (!raised) { YP_CUDA_OK(cudaFuncSetAttribute(yp::kp_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024)); raised = true; }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem; cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  YP_CUDA_OK(cudaLaunchKernelEx(&cfg, yp::kp_filter_kernel, pts_all, n_all, max_pts, H, W, boxes, box_count, box_ld, out_pts, out_sel, out_count,
                                n_prev, row_key, col_key));
  return YP_OK;
}

extern "C" int yp_keypoints_threshold_count(const void* workspace, size_t workspace_bytes, int32_t B, int32_t H, int32_t W, int32_t max_pts,
                                            int32_t* out_count, void* stream) {
  YP_REQUIRE(workspace && out_count, YP_ERR_ARG, "keypoints_threshold_count: null pointer");
  yp::KpWs ws;
  const size_t need = yp::carve(&ws, static_cast<char*>(const_cast<void*>(workspace)), B, H, W, max_pts);
  YP_REQUIRE(workspace_bytes >= need, YP_ERR_CAPACITY, "keypoints_threshold_count: workspace %zu < %zu bytes", workspace_bytes, need);
  YP_CUDA_OK(cudaMemcpyAsync(out_count, ws.n_thresh, sizeof(int) * B, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return YP_OK;
}
