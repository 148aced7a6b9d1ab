// Detect decode + box NMS (models/yolo.py:49-81, utils/general_yolo.py:124-235 of the reference).
// HBM/L2-bound integer+fp32 work; every fp32 expression that feeds a comparison is written with explicit
// round-to-nearest intrinsics in the reference's operation order so that keep/suppress decisions and
// indices are bit-identical to the PyTorch/torchvision CPU path (no FMA contraction).
#include <algorithm>

#include "common.cuh"

namespace yp {
namespace {

// ------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------
struct DecodeArgs {
  const float* logits;
  float* raw;
  float* pred;
  int B, ny, nx, ldc, na, no;
  float stride;
  float anchor[16];
  long long A_total, row_off;
};

// One CTA owns `pix_per_cta` consecutive pixels of the flattened [B, ny, nx] space; thread t owns channel t (= anchor an, output
// o) and walks the pixels, so (an, o, anchor size) are computed once per thread and the loop body has no integer division or
// 64-bit multiply: consecutive pixels of one image are consecutive rows of `raw` / `pred` for a fixed anchor, so both output
// pointers advance by `no` floats per pixel and are only recomputed when the walk crosses into the next image.  One coalesced
// 4-byte load (channels are the fastest logits dimension) and two stores per element into the 340-byte (b, an, y, x) rows, whose
// neighbours x+1 follow in the next iteration of the same CTA (L2 merges the partial sectors at the row ends).  The first version
// did four 64-bit divisions per element and reached 16-21 % of the HBM peak.
constexpr int kDecodeBatch = 8;

__global__ void __launch_bounds__(256) detect_decode_kernel(const DecodeArgs a, int pix_per_cta) {
  const int nch = a.na * a.no;
  const int64_t npix = static_cast<int64_t>(a.B) * a.ny * a.nx;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * pix_per_cta;
  const int n = static_cast<int>(min(static_cast<int64_t>(pix_per_cta), npix - p0));
  const int x0 = static_cast<int>(p0 % a.nx);
  const int64_t t0 = p0 / a.nx;
  const int y0 = static_cast<int>(t0 % a.ny), b0 = static_cast<int>(t0 / a.ny);
  const int64_t plane = static_cast<int64_t>(a.ny) * a.nx;
  for (int ch = threadIdx.x; ch < nch; ch += blockDim.x) {
    const int an = ch / a.no, o = ch - an * a.no;
    const float anc = (o == 2 || o == 3) ? a.anchor[an * 2 + (o - 2)] : 0.0f;
    const float* src = a.logits + p0 * a.ldc + ch;
    int x = x0, y = y0, b = b0;
    auto raw_at = [&](int bb, int yy, int xx) { return a.raw ? a.raw + ((static_cast<int64_t>(bb) * a.na + an) * plane + static_cast<int64_t>(yy) * a.nx + xx) * a.no + o : nullptr; };
    auto pred_at = [&](int bb, int yy, int xx) { return a.pred + (static_cast<int64_t>(bb) * a.A_total + a.row_off + an * plane + static_cast<int64_t>(yy) * a.nx + xx) * a.no + o; };
    float* rp = raw_at(b, y, x);
    float* pp = pred_at(b, y, x);
    float xf = static_cast<float>(x), yf = static_cast<float>(y);      // float copies of the grid position (no int -> float conversion per element)
    // Batches of kDecodeBatch independent loads are issued before the first use: with the load inside the (branchy) per-element body
    // every iteration waited a full DRAM round trip (ncu: all stall samples on the first instruction of expf).
    for (int i0 = 0; i0 < n; i0 += kDecodeBatch) {
      float v[kDecodeBatch];
#pragma unroll
      for (int u = 0; u < kDecodeBatch; ++u) v[u] = i0 + u < n ? __ldg(src + static_cast<int64_t>(i0 + u) * a.ldc) : 0.0f;
#pragma unroll
      for (int u = 0; u < kDecodeBatch; ++u) {
        if (i0 + u < n) {
          if (rp) { *rp = v[u]; rp += a.no; }
          const float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v[u])));
          const float t = __fmul_rn(s, 2.0f);
          const float xy = __fmul_rn(__fadd_rn(__fsub_rn(t, 0.5f), o == 0 ? xf : yf), a.stride);
          const float wh = __fmul_rn(__fmul_rn(t, t), anc);
          *pp = o < 2 ? xy : (o < 4 ? wh : s);
          pp += a.no;
          xf += 1.0f;
          if (++x == a.nx) {
            x = 0; xf = 0.0f; yf += 1.0f;
            if (++y == a.ny) { y = 0; yf = 0.0f; ++b; rp = raw_at(b, 0, 0); pp = pred_at(b, 0, 0); }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// NMS stage 1: candidates.  One warp per prediction row.
// ------------------------------------------------------------------------------------------------
struct NmsWs {
  int* n_cand;          // [B]  candidates found (may exceed cap); slots are handed out with atomicAdd
  int* n_sorted;        // [B]  min(n_cand, cap, max_nms)
  unsigned int* ord;    // [B][cap]  position of the candidate in the reference's row-major (box, class) order
  float* cand;          // [B][cap][6]  x1,y1,x2,y2,conf,cls in arbitrary (slot) order
  float* sorted;        // [B][cap][6]  confidence-descending (stable)
  unsigned long long* mask;  // [B][cap][cap/64]
};

__device__ __forceinline__ bool class_ok(const uint32_t* class_mask, int c) {
  return class_mask == nullptr || ((class_mask[c >> 5] >> (c & 31)) & 1u);
}

// Candidate emission shared by the two front ends (decoded `pred` rows, or raw Detect logits).
// x,y,w,h,obj and a class-score accessor cls(c) are already sigmoid-decoded values.
template <typename ClsFn>
__device__ __forceinline__ void emit_candidates(float bx, float by, float bw, float bh, float obj, int nc, ClsFn cls, unsigned int row,
                                                const YpNmsParams& p, int cap, int b, const NmsWs& ws) {
  auto put = [&](float conf, int c, unsigned int ord) {
    const int slot = atomicAdd(&ws.n_cand[b], 1);
    if (slot < cap) {
      float* o = ws.cand + (static_cast<int64_t>(b) * cap + slot) * 6;
      o[0] = __fsub_rn(bx, __fdiv_rn(bw, 2.0f)); o[1] = __fsub_rn(by, __fdiv_rn(bh, 2.0f));   // xywh2xyxy, general_yolo.py:623-630
      o[2] = __fadd_rn(bx, __fdiv_rn(bw, 2.0f)); o[3] = __fadd_rn(by, __fdiv_rn(bh, 2.0f));
      o[4] = conf; o[5] = static_cast<float>(c);
      ws.ord[static_cast<int64_t>(b) * cap + slot] = ord;
    }
  };
  if (p.multi_label && nc > 1) {  // one candidate per (row, class) with conf > thr, general_yolo.py:191-193
    for (int c = 0; c < nc; ++c) {
      const float conf = __fmul_rn(cls(c), obj);
      if (conf > p.conf_thres && class_ok(p.class_mask, c)) put(conf, c, row * static_cast<unsigned int>(nc) + c);
    }
  } else {                        // best class only (first maximum), general_yolo.py:195-196
    float best = -INFINITY;
    int bc = 0;
    for (int c = 0; c < nc; ++c) {
      const float conf = __fmul_rn(cls(c), obj);
      if (conf > best) { best = conf; bc = c; }
    }
    if (best > p.conf_thres && class_ok(p.class_mask, bc)) put(best, bc, row);
  }
}

// front end 1: decoded predictions [B,A,no]; one thread per row (objectness survivors are rare)
__global__ void nms_candidates_pred_kernel(const float* __restrict__ pred, int B, long long A, int no, YpNmsParams p, int cap, NmsWs ws) {
  const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (row >= static_cast<int64_t>(B) * A) return;
  const int b = static_cast<int>(row / A);
  const float* r = pred + row * no;
  const float obj = r[4];
  if (!(obj > p.conf_thres)) return;  // strict, general_yolo.py:146
  emit_candidates(r[0], r[1], r[2], r[3], obj, no - 5, [&](int c) { return r[5 + c]; }, static_cast<unsigned int>(row - b * A), p, cap, b, ws);
}

// front end 2: raw Detect logits of the three levels (NHWC, channel a*no+o): decode (models/yolo.py:60-68) only the rows
// whose objectness passes, so `pred` is never materialised in the whole-frame pipeline.
struct DetLevels {
  const float* logits[3];
  int ny[3], nx[3], ldc[3];
  float stride[3];
  float anchor[3][6];
  long long row_off[3];
  int na, no;
};

__device__ __forceinline__ float sigmoid_rn(float v) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v))); }

__global__ void nms_candidates_logits_kernel(const DetLevels lv, int B, long long A, YpNmsParams p, int cap, NmsWs ws) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * A) return;
  const int b = static_cast<int>(idx / A);
  const long long row = idx - b * A;
  const int l = row >= lv.row_off[2] ? 2 : (row >= lv.row_off[1] ? 1 : 0);
  const int ny = lv.ny[l], nx = lv.nx[l];
  long long cell = row - lv.row_off[l];                 // (anchor, y, x) order, models/yolo.py:56
  const int x = static_cast<int>(cell % nx); cell /= nx;
  const int y = static_cast<int>(cell % ny);
  const int an = static_cast<int>(cell / ny);
  const float* r = lv.logits[l] + ((static_cast<int64_t>(b) * ny + y) * nx + x) * lv.ldc[l] + an * lv.no;
  const float obj = sigmoid_rn(r[4]);
  if (!(obj > p.conf_thres)) return;
  const float sx = sigmoid_rn(r[0]), sy = sigmoid_rn(r[1]), sw = sigmoid_rn(r[2]), sh = sigmoid_rn(r[3]);
  const float bx = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sx, 2.0f), 0.5f), static_cast<float>(x)), lv.stride[l]);
  const float by = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sy, 2.0f), 0.5f), static_cast<float>(y)), lv.stride[l]);
  const float tw = __fmul_rn(sw, 2.0f), th = __fmul_rn(sh, 2.0f);
  const float bw = __fmul_rn(__fmul_rn(tw, tw), lv.anchor[l][an * 2]);
  const float bh = __fmul_rn(__fmul_rn(th, th), lv.anchor[l][an * 2 + 1]);
  emit_candidates(bx, by, bw, bh, obj, lv.no - 5, [&](int c) { return sigmoid_rn(r[5 + c]); }, static_cast<unsigned int>(row), p, cap, b, ws);
}

__global__ void nms_counts_kernel(int B, int cap, int max_nms, NmsWs ws) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = min(ws.n_cand[b], cap);
  ws.n_sorted[b] = min(n, max_nms);
}

// descending rank sort, ties in the reference's candidate order: rank(i) = #{j : conf_j > conf_i or (conf_j == conf_i and ord_j < ord_i)}
__global__ void nms_rank_kernel(int cap, NmsWs ws) {
  __shared__ float tile[256];
  __shared__ unsigned int tord[256];
  const int b = blockIdx.y;
  const int n = min(ws.n_cand[b], cap);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x * blockDim.x >= n) return;
  const float* cand = ws.cand + static_cast<int64_t>(b) * cap * 6;
  const unsigned int* ord = ws.ord + static_cast<int64_t>(b) * cap;
  const float ci = i < n ? cand[i * 6 + 4] : 0.0f;
  const unsigned int oi = i < n ? ord[i] : 0u;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    const int j = j0 + threadIdx.x;
    tile[threadIdx.x] = j < n ? cand[j * 6 + 4] : -INFINITY;
    tord[threadIdx.x] = j < n ? ord[j] : 0xffffffffu;
    __syncthreads();
    const int lim = min(256, n - j0);
    for (int k = 0; k < lim; ++k) {
      const float cj = tile[k];
      rank += (cj > ci || (cj == ci && tord[k] < oi)) ? 1 : 0;
    }
    __syncthreads();
  }
  if (i < n && rank < ws.n_sorted[b]) {
    float* o = ws.sorted + (static_cast<int64_t>(b) * cap + rank) * 6;
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = cand[i * 6 + k];
  }
}

// torchvision nms_kernel semantics: suppress j (> i in sorted order) iff inter / (area_i + area_j - inter) > thr
__device__ __forceinline__ bool iou_gt(const float* a, const float* b, float thr) {
  const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  const float w = fmaxf(__fsub_rn(right, left), 0.0f), h = fmaxf(__fsub_rn(bottom, top), 0.0f);
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]));
  const float sb = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > thr;
}

// 64x64 blocks of the upper-triangular suppression bit matrix; persistent over block pairs
__global__ void nms_mask_kernel(int cap, float iou_thres, int agnostic, float max_wh, NmsWs ws) {
  __shared__ float colbox[64][4];
  const int b = blockIdx.y;
  const int n = ws.n_sorted[b];
  const int nb = (n + 63) >> 6;
  const int words = cap >> 6;
  const float* sorted = ws.sorted + static_cast<int64_t>(b) * cap * 6;
  unsigned long long* mask = ws.mask + static_cast<int64_t>(b) * cap * words;
  for (int t = blockIdx.x; t < nb * nb; t += gridDim.x) {
    const int rb = t / nb, cb = t - rb * nb;
    if (cb < rb) continue;
    __syncthreads();
    {
      const int j = cb * 64 + threadIdx.x;
      if (j < n) {
        const float off = agnostic ? 0.0f : __fmul_rn(sorted[j * 6 + 5], max_wh);
#pragma unroll
        for (int k = 0; k < 4; ++k) colbox[threadIdx.x][k] = __fadd_rn(sorted[j * 6 + k], off);
      }
    }
    __syncthreads();
    const int i = rb * 64 + threadIdx.x;
    if (i < n) {
      float me[4];
      const float off = agnostic ? 0.0f : __fmul_rn(sorted[i * 6 + 5], max_wh);
#pragma unroll
      for (int k = 0; k < 4; ++k) me[k] = __fadd_rn(sorted[i * 6 + k], off);
      unsigned long long bits = 0;
      const int lim = min(64, n - cb * 64);
      const int start = (rb == cb) ? threadIdx.x + 1 : 0;
      for (int k = start; k < lim; ++k)
        if (iou_gt(me, colbox[k], iou_thres)) bits |= 1ull << k;
      mask[static_cast<int64_t>(i) * words + cb] = bits;
    }
  }
}

// ordered scan over the bit matrix; one block per image.  The upper-triangular part of the matrix that the scan touches
// (n rows x ceil(n/64) words) is staged in shared memory when it fits, the 64-step dependent scan of a chunk runs in one
// warp on registers (diagonal words exchanged by shuffles), and the suppression rows of the kept boxes are folded in
// parallel.
constexpr int kScanSmemBytes = 160 * 1024;

__global__ void __launch_bounds__(256) nms_scan_keep_kernel(int cap, int max_det, NmsWs ws, float* __restrict__ out_boxes, int* __restrict__ out_count,
                                                            int stage_words) {
  extern __shared__ unsigned long long scan_smem[];   // removed[cap/64] then (optionally) the staged matrix
  __shared__ unsigned long long keep_bits;
  __shared__ int n_keep;
  const int b = blockIdx.x;
  const int n = ws.n_sorted[b];
  const int nb = (n + 63) >> 6;
  const int words = cap >> 6;
  unsigned long long* removed = scan_smem;
  unsigned long long* staged = scan_smem + words;
  const bool in_smem = static_cast<long long>(n) * nb <= stage_words;
  const float* sorted = ws.sorted + static_cast<int64_t>(b) * cap * 6;
  const unsigned long long* mask = ws.mask + static_cast<int64_t>(b) * cap * words;
  float* out = out_boxes + static_cast<int64_t>(b) * max_det * 6;
  for (int w = threadIdx.x; w < nb; w += blockDim.x) removed[w] = 0;
  if (in_smem)
    for (int t = threadIdx.x; t < n * nb; t += blockDim.x) {
      const int i = t / nb, w = t - i * nb;
      staged[t] = w >= (i >> 6) ? mask[static_cast<int64_t>(i) * words + w] : 0ull;   // words left of the diagonal are never written
    }
  if (threadIdx.x == 0) n_keep = 0;
  __syncthreads();
  auto row_word = [&](int i, int w) -> unsigned long long {
    return in_smem ? staged[i * nb + w] : mask[static_cast<int64_t>(i) * words + w];
  };
  for (int wc = 0; wc < nb; ++wc) {
    if (n_keep >= max_det) break;  // uniform: n_keep only changes between barriers
    const int lim = min(64, n - wc * 64);
    const int base = n_keep;
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      const unsigned long long d0 = lane < lim ? row_word(wc * 64 + lane, wc) : 0ull;
      const unsigned long long d1 = lane + 32 < lim ? row_word(wc * 64 + 32 + lane, wc) : 0ull;
      unsigned long long rem = removed[wc], kb = 0;
      int nk = base;
      for (int k = 0; k < lim && nk < max_det; ++k) {
        const unsigned long long dk = __shfl_sync(0xffffffffu, k < 32 ? d0 : d1, k & 31);
        if (!((rem >> k) & 1ull)) { kb |= 1ull << k; rem |= dk; ++nk; }
      }
      if (lane == 0) keep_bits = kb;
    }
    __syncthreads();
    const unsigned long long kb = keep_bits;
    if (threadIdx.x < 64 && ((kb >> threadIdx.x) & 1ull)) {   // kept rows, in order
      const int pos = base + __popcll(kb & ((1ull << threadIdx.x) - 1ull));
      const float* sr = sorted + static_cast<int64_t>(wc * 64 + threadIdx.x) * 6;
#pragma unroll
      for (int k = 0; k < 6; ++k) out[pos * 6 + k] = sr[k];
    }
    // fold the suppression rows of the kept boxes into `removed`: one (row, word) pair per thread
    const int later = nb - (wc + 1);
    for (int t = threadIdx.x; t < 64 * later; t += blockDim.x) {
      const int k = t / later, w = wc + 1 + (t - k * later);
      if ((kb >> k) & 1ull) {
        const unsigned long long m = row_word(wc * 64 + k, w);
        if (m) atomicOr(&removed[w], m);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) n_keep = base + __popcll(kb);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int nc = ws.n_cand[b];
    out_count[b] = nc > cap ? -1 - nc : n_keep;
  }
}

size_t align_up(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

size_t carve(NmsWs* ws, char* base, int B, long long A, int cap) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
  (void)A;
  ws->n_cand = reinterpret_cast<int*>(take(sizeof(int) * B));
  ws->n_sorted = reinterpret_cast<int*>(take(sizeof(int) * B));
  ws->ord = reinterpret_cast<unsigned int*>(take(sizeof(unsigned int) * B * cap));
  ws->cand = reinterpret_cast<float*>(take(sizeof(float) * 6 * B * cap));
  ws->sorted = reinterpret_cast<float*>(take(sizeof(float) * 6 * B * cap));
  ws->mask = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * B * cap * (cap / 64)));
  return off;
}

int nms_tail(int B, int cap, const YpNmsParams& p, const NmsWs& ws, float* out_boxes, int32_t* out_count, cudaStream_t st) {
  nms_counts_kernel<<<ceil_div(B, 128), 128, 0, st>>>(B, cap, p.max_nms, ws);
  nms_rank_kernel<<<dim3(cap / 256 + (cap % 256 ? 1 : 0), B), 256, 0, st>>>(cap, ws);
  nms_mask_kernel<<<dim3(2 * sm_count(), B), 64, 0, st>>>(cap, p.iou_thres, p.agnostic, p.max_wh, ws);
  static thread_local bool raised = false;
  if (!raised) { YP_CUDA_OK(cudaFuncSetAttribute(nms_scan_keep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kScanSmemBytes)); raised = true; }
  const int stage_words = (kScanSmemBytes - static_cast<int>(sizeof(unsigned long long)) * (cap / 64)) / static_cast<int>(sizeof(unsigned long long));
  nms_scan_keep_kernel<<<B, 256, kScanSmemBytes, st>>>(cap, p.max_det, ws, out_boxes, out_count, stage_words > 0 ? stage_words : 0);
  YP_LAUNCH_OK();
  return YP_OK;
}

}  // namespace
}  // namespace yp

extern "C" int yp_detect_decode(const float* logits, int32_t B, int32_t ny, int32_t nx, int32_t ldc, int32_t na, int32_t no,
                                float stride_px, const float* anchors_px_host, float* raw, float* pred, int64_t A_total,
                                int64_t row_off, void* stream) {
  YP_REQUIRE(logits && pred && anchors_px_host, YP_ERR_ARG, "detect_decode: null pointer");
  YP_REQUIRE(na >= 1 && na <= 8 && no >= 6 && na * no <= ldc, YP_ERR_SHAPE, "detect_decode: na=%d no=%d ldc=%d", na, no, ldc);
  yp::DecodeArgs a;
  a.logits = logits; a.raw = raw; a.pred = pred; a.B = B; a.ny = ny; a.nx = nx; a.ldc = ldc; a.na = na; a.no = no;
  a.stride = stride_px; a.A_total = A_total; a.row_off = row_off;
  for (int i = 0; i < na * 2; ++i) a.anchor[i] = anchors_px_host[i];
  YP_REQUIRE(B > 0 && ny > 0 && nx > 0, YP_ERR_SHAPE, "detect_decode: B=%d ny=%d nx=%d", B, ny, nx);
  const int64_t npix = static_cast<int64_t>(B) * ny * nx;
  const int pix_per_cta = 32;
  const int threads = std::min(256, (na * no + 31) / 32 * 32);
  yp::detect_decode_kernel<<<static_cast<unsigned>(yp::ceil_div64(npix, pix_per_cta)), threads, 0, static_cast<cudaStream_t>(stream)>>>(a, pix_per_cta);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" size_t yp_box_nms_workspace_bytes(int32_t B, int64_t A, int32_t no, int32_t cap) {
  (void)no;
  if (B <= 0 || A <= 0 || cap <= 0 || cap % 64) return 0;
  yp::NmsWs ws;
  return yp::carve(&ws, nullptr, B, A, cap);
}

extern "C" int yp_box_nms(const float* pred, int32_t B, int64_t A, int32_t no, const YpNmsParams* p, int32_t cap,
                          float* out_boxes, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
  YP_REQUIRE(pred && p && out_boxes && out_count && workspace, YP_ERR_ARG, "box_nms: null pointer");
  YP_REQUIRE(B > 0 && A > 0 && no >= 6, YP_ERR_SHAPE, "box_nms: B=%d A=%lld no=%d", B, (long long)A, no);
  YP_REQUIRE(cap > 0 && cap % 64 == 0, YP_ERR_SHAPE, "box_nms: cap=%d must be a positive multiple of 64", cap);
  YP_REQUIRE(p->conf_thres >= 0.f && p->conf_thres <= 1.f, YP_ERR_ARG, "Invalid Confidence threshold %g, valid values are between 0.0 and 1.0", p->conf_thres);
  YP_REQUIRE(p->iou_thres >= 0.f && p->iou_thres <= 1.f, YP_ERR_ARG, "Invalid IoU %g, valid values are between 0.0 and 1.0", p->iou_thres);
  YP_REQUIRE(p->max_det > 0 && p->max_nms > 0, YP_ERR_ARG, "box_nms: max_det/max_nms must be positive");
  yp::NmsWs ws;
  const size_t need = yp::carve(&ws, static_cast<char*>(workspace), B, A, cap);
  YP_REQUIRE(workspace_bytes >= need, YP_ERR_CAPACITY, "box_nms: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t rows = static_cast<int64_t>(B) * A;
  YP_CUDA_OK(cudaMemsetAsync(ws.n_cand, 0, sizeof(int) * B, st));
  yp::nms_candidates_pred_kernel<<<static_cast<unsigned>(yp::ceil_div64(rows, 256)), 256, 0, st>>>(pred, B, A, no, *p, cap, ws);
  return yp::nms_tail(B, cap, *p, ws, out_boxes, out_count, st);
}

extern "C" int yp_detect_nms(const float* const* logits3, const int32_t* ny3, const int32_t* nx3, const int32_t* ldc3, const float* stride3,
                             const float* anchors_px_host, int32_t B, int32_t na, int32_t no, const YpNmsParams* p, int32_t cap,
                             float* out_boxes, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
  YP_REQUIRE(logits3 && ny3 && nx3 && ldc3 && stride3 && anchors_px_host && p && out_boxes && out_count && workspace, YP_ERR_ARG, "detect_nms: null pointer");
  YP_REQUIRE(B > 0 && na == 3 && no >= 6, YP_ERR_SHAPE, "detect_nms: B=%d na=%d no=%d (na must be 3)", B, na, no);
  YP_REQUIRE(cap > 0 && cap % 64 == 0, YP_ERR_SHAPE, "detect_nms: cap=%d must be a positive multiple of 64", cap);
  YP_REQUIRE(p->conf_thres >= 0.f && p->conf_thres <= 1.f, YP_ERR_ARG, "Invalid Confidence threshold %g, valid values are between 0.0 and 1.0", p->conf_thres);
  YP_REQUIRE(p->iou_thres >= 0.f && p->iou_thres <= 1.f, YP_ERR_ARG, "Invalid IoU %g, valid values are between 0.0 and 1.0", p->iou_thres);
  YP_REQUIRE(p->max_det > 0 && p->max_nms > 0, YP_ERR_ARG, "detect_nms: max_det/max_nms must be positive");
  yp::DetLevels lv;
  long long A = 0;
  for (int l = 0; l < 3; ++l) {
    YP_REQUIRE(logits3[l] && na * no <= ldc3[l], YP_ERR_SHAPE, "detect_nms: level %d logits missing or ldc too small", l);
    lv.logits[l] = logits3[l]; lv.ny[l] = ny3[l]; lv.nx[l] = nx3[l]; lv.ldc[l] = ldc3[l]; lv.stride[l] = stride3[l];
    for (int i = 0; i < 6; ++i) lv.anchor[l][i] = anchors_px_host[l * 6 + i];
    lv.row_off[l] = A;
    A += static_cast<long long>(na) * ny3[l] * nx3[l];
  }
  lv.na = na; lv.no = no;
  yp::NmsWs ws;
  const size_t need = yp::carve(&ws, static_cast<char*>(workspace), B, A, cap);
  YP_REQUIRE(workspace_bytes >= need, YP_ERR_CAPACITY, "detect_nms: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  YP_CUDA_OK(cudaMemsetAsync(ws.n_cand, 0, sizeof(int) * B, st));
  yp::nms_candidates_logits_kernel<<<static_cast<unsigned>(yp::ceil_div64(static_cast<int64_t>(B) * A, 256)), 256, 0, st>>>(lv, B, A, *p, cap, ws);
  return yp::nms_tail(B, cap, *p, ws, out_boxes, out_count, st);
}


