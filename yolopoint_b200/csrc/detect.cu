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
// Box NMS: ONE kernel, one 1024-thread CTA per image (general_yolo.py:124-235 + torchvision.ops.nms).
//
//   A  scan the objectness of all A rows, list the rows with obj > conf_thres            (general_yolo.py:146, 168)
//   B  evaluate conf = cls * obj of the listed rows, one (row, class) pair per thread step; every candidate is ONE 64-bit key
//      (conf bits << 32 | ~ord), ord = row * nc + class = the reference's row-major (box, class) order   (:184-196)
//   C  more candidates than the buffer holds (cap >= max_nms): exact radix select of the max_nms largest keys over re-scans of
//      the listed rows -- the reference's `argsort(descending=True)[:max_nms]` (:210-211), never an error
//   D  bitonic sort of the keys, descending = confidence descending, ties in candidate order  (:211-213)
//   E  boxes of the sorted candidates are re-derived from their rows (+ class offset unless agnostic, :216-217)
//   F  greedy suppression in chunks of 64 sorted boxes: the chunk's 64 x 64 IoU bits, a sequential resolve of the chunk by one
//      thread, then the chunk's kept boxes clear all later boxes in parallel.  IoU work is kept x n instead of n^2 / 2 and no
//      n x n bit matrix exists, so 30 000 candidates need 1 MB of scratch, not 113 MB.
//
// Up to kSmemCap candidates everything lives in shared memory (the usual frame: 10^2..10^3 candidates, ~10 us); beyond that the
// same code runs on arrays in the caller's workspace (L2).  The five-kernel pipeline this replaces (candidates / counts / rank /
// mask / scan) cost 77 us of launch + dependency latency per frame and aborted above `cap` candidates.
// ------------------------------------------------------------------------------------------------
constexpr int kNmsThreads = 1024;
constexpr int kSmemCap = 4096;      // candidates whose keys + boxes fit in shared memory
constexpr int kPassSmem = 4096;     // objectness survivors listed in shared memory (the rest spill to the workspace)

struct NmsWs {
  int* stats;                  // [B][4]  rows passing objectness, candidates found, candidates sorted, path (0 smem, 1 workspace, 2 select)
  unsigned long long* pass_row; // [B][A]  row handles
  float* pass_obj;             // [B][A]
  unsigned long long* key;     // [B][cap_p2]
  float4* box;                 // [B][cap]   NMS boxes (class offset applied) in sorted order
  unsigned long long* removed; // [B][cap_p2 / 64]
  int* pre_n;                  // [B][2]  prescan: candidates emitted / true number of (row, class) pairs; zero between frames
  unsigned long long* pre_key; // [B][kSmemCap]  prescan: candidate keys of the levels scanned ahead of the NMS kernel
};

__device__ __forceinline__ bool class_ok(const uint32_t* class_mask, int c) {
  return class_mask == nullptr || ((class_mask[c >> 5] >> (c & 31)) & 1u);
}

__device__ __forceinline__ float sigmoid_rn(float v) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v))); }

// xywh -> xyxy exactly as general_yolo.py:623-630
__device__ __forceinline__ float4 xywh2xyxy_rn(float bx, float by, float bw, float bh) {
  return make_float4(__fsub_rn(bx, __fdiv_rn(bw, 2.0f)), __fsub_rn(by, __fdiv_rn(bh, 2.0f)), __fadd_rn(bx, __fdiv_rn(bw, 2.0f)),
                     __fadd_rn(by, __fdiv_rn(bh, 2.0f)));
}

// A front end enumerates the rows of one image as (segment, position) pairs without integer division by run-time values, hands
// out a 64-bit row handle (what the later phases need to find the row again) and answers three questions about a row:
// objectness, class score c, decoded box.  `*_may_pass` are cheap conservative pre-tests on the raw value (no false negatives):
// the exact test `value > thr` is only evaluated where they hold.
//
// front end 1: decoded predictions [B, A, no]
struct PredRows {
  const float* pred;
  long long A;
  int no;
  float thr;
  __device__ __forceinline__ int segments() const { return 1; }
  __device__ __forceinline__ int seg_rows(int) const { return static_cast<int>(A); }
  __device__ __forceinline__ unsigned long long handle(int, int i) const { return static_cast<unsigned long long>(i); }
  __device__ __forceinline__ unsigned int row_of(unsigned long long h) const { return static_cast<unsigned int>(h); }
  __device__ __forceinline__ unsigned long long handle_of_row(unsigned int r) const { return r; }
  __device__ __forceinline__ const float* row(int b, unsigned long long h) const { return pred + (static_cast<int64_t>(b) * A + static_cast<int64_t>(h)) * no; }
  __device__ __forceinline__ float obj_raw(int b, unsigned long long h) const { return __ldg(row(b, h) + 4); }
  __device__ __forceinline__ bool obj_may_pass(float raw) const { return raw > thr; }
  __device__ __forceinline__ float obj(float raw) const { return raw; }
  __device__ __forceinline__ float cls_raw(int b, unsigned long long h, int c) const { return __ldg(row(b, h) + 5 + c); }
  __device__ __forceinline__ bool cls_may_pass(float raw, float) const { return true; }
  __device__ __forceinline__ float cls(float raw) const { return raw; }
  __device__ __forceinline__ float4 box(int b, unsigned long long h) const {
    const float* p = row(b, h);
    return xywh2xyxy_rn(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
  }
};

// front end 2: raw Detect logits of the three levels (NHWC, channel a*no+o): decode (models/yolo.py:60-68) only what the NMS
// looks at, so `pred` is never materialised in the whole-frame pipeline.  Segment = level, position = pixel * 3 + anchor
// (consecutive threads read the three objectness logits of consecutive pixels); handle = level | anchor | y | x | row.
struct DetLevels {
  const float* logits[3];
  int ny[3], nx[3], ldc[3];
  float stride[3];
  float anchor[3][6];
  long long row_off[3];
  long long A;
  int na, no;
  float logit_thr;   // log(thr / (1 - thr)) - margin: sigmoid(v) > thr implies v > logit_thr
  __device__ __forceinline__ int segments() const { return 3; }
  __device__ __forceinline__ int seg_rows(int l) const { return 3 * ny[l] * nx[l]; }
  __device__ __forceinline__ unsigned long long handle(int l, int i) const {
    const int pix = i / 3, an = i - pix * 3;
    const int y = pix / nx[l], x = pix - y * nx[l];
    const unsigned int r = static_cast<unsigned int>(row_off[l]) + static_cast<unsigned int>((an * ny[l] + y) * nx[l] + x);   // models/yolo.py:56
    return (static_cast<unsigned long long>(l) << 62) | (static_cast<unsigned long long>(an) << 60) | (static_cast<unsigned long long>(y) << 46) |
           (static_cast<unsigned long long>(x) << 32) | r;
  }
  __device__ __forceinline__ unsigned int row_of(unsigned long long h) const { return static_cast<unsigned int>(h); }
  __device__ __forceinline__ unsigned long long handle_of_row(unsigned int r) const {
    const int l = r >= row_off[2] ? 2 : (r >= row_off[1] ? 1 : 0);
    unsigned int cell = r - static_cast<unsigned int>(row_off[l]);
    const unsigned int x = cell % nx[l]; cell /= nx[l];
    const unsigned int y = cell % ny[l], an = cell / ny[l];
    return (static_cast<unsigned long long>(l) << 62) | (static_cast<unsigned long long>(an) << 60) | (static_cast<unsigned long long>(y) << 46) |
           (static_cast<unsigned long long>(x) << 32) | r;
  }
  __device__ __forceinline__ const float* row(int b, unsigned long long h) const {
    const int l = static_cast<int>(h >> 62), an = static_cast<int>(h >> 60) & 3, y = static_cast<int>(h >> 46) & 0x3fff, x = static_cast<int>(h >> 32) & 0x3fff;
    return logits[l] + ((static_cast<int64_t>(b) * ny[l] + y) * nx[l] + x) * ldc[l] + an * no;
  }
  // position i of segment l without building the handle: pixel-major NHWC, 3 anchors per pixel
  __device__ __forceinline__ float obj_raw_at(int b, int l, int i) const {
    const int pix = i / 3, an = i - pix * 3;
    return __ldg(logits[l] + (static_cast<int64_t>(b) * ny[l] * nx[l] + pix) * ldc[l] + an * no + 4);
  }
  __device__ __forceinline__ bool obj_may_pass(float raw) const { return raw > logit_thr; }
  __device__ __forceinline__ float obj(float raw) const { return sigmoid_rn(raw); }
  __device__ __forceinline__ float cls_raw(int b, unsigned long long h, int c) const { return __ldg(row(b, h) + 5 + c); }
  __device__ __forceinline__ bool cls_may_pass(float raw, float) const { return raw > logit_thr; }   // conf = cls * obj <= cls
  __device__ __forceinline__ float cls(float raw) const { return sigmoid_rn(raw); }
  __device__ __forceinline__ float4 box(int b, unsigned long long h) const {
    const int l = static_cast<int>(h >> 62), an = static_cast<int>(h >> 60) & 3, y = static_cast<int>(h >> 46) & 0x3fff, x = static_cast<int>(h >> 32) & 0x3fff;
    const float* p = row(b, h);
    const float sx = sigmoid_rn(__ldg(p)), sy = sigmoid_rn(__ldg(p + 1)), sw = sigmoid_rn(__ldg(p + 2)), sh = sigmoid_rn(__ldg(p + 3));
    const float bx = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sx, 2.0f), 0.5f), static_cast<float>(x)), stride[l]);
    const float by = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sy, 2.0f), 0.5f), static_cast<float>(y)), stride[l]);
    const float tw = __fmul_rn(sw, 2.0f), th = __fmul_rn(sh, 2.0f);
    return xywh2xyxy_rn(bx, by, __fmul_rn(__fmul_rn(tw, tw), anchor[l][an * 2]), __fmul_rn(__fmul_rn(th, th), anchor[l][an * 2 + 1]));
  }
};
template <class FE> __device__ __forceinline__ float fe_obj_raw_at(const FE& fe, int b, int l, int i) { return fe.obj_raw(b, fe.handle(l, i)); }
template <> __device__ __forceinline__ float fe_obj_raw_at<DetLevels>(const DetLevels& fe, int b, int l, int i) { return fe.obj_raw_at(b, l, i); }

// torchvision nms_kernel semantics: suppress j (after i in sorted order) iff inter / (area_i + area_j - inter) > thr.
// inter == 0 can never exceed thr >= 0 (0 / x is 0, -0 or NaN), so disjoint pairs skip the division.
__device__ __forceinline__ bool iou_gt(const float4& a, const float4& b, float thr) {
  const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
  const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
  if (!(w > 0.0f && h > 0.0f)) return false;
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > thr;
}

__device__ __forceinline__ unsigned long long cand_key(float conf, unsigned int ord) {
  return (static_cast<unsigned long long>(__float_as_uint(conf)) << 32) | (0xffffffffu - ord);   // conf > 0: bits are monotone
}


// The candidates of ONE objectness survivor, evaluated by one warp (lanes over the classes: coalesced logits).  f(key) is called
// once per candidate; returns (to every lane) how many (row, class) pairs exceed the threshold.
//
// `dedupe` (agnostic multi-label NMS, iou_thres < 1): the candidates of one row share one box, so their mutual IoU is exactly 1
// (inter == area_i == area_j bit for bit, (a + a) - a == a in fp32) and only the most confident one can ever be kept: it
// precedes the others in the sort, suppresses them if it is kept, and whatever suppresses it (a kept box of higher confidence
// with IoU > thr) suppresses them too -- same box, same IoU.  Emitting only that candidate leaves the output unchanged and cuts
// the sort / suppression work by the number of classes per row (5.5 on the benchmark frames).  Boxes without a finite positive
// area are exempt (their IoU is 0 or NaN, nothing is suppressed).  The caller must fall back to all candidates when their true
// number exceeds max_nms (the reference's top-max_nms cut is taken over all of them).
template <class FE, class F>
__device__ __forceinline__ int warp_row_candidates(const FE& fe, int b, unsigned long long h, float obj, const YpNmsParams& p, int nc,
                                                   bool multi, bool dedupe, int lane, F&& f) {
  const unsigned int ord0 = fe.row_of(h) * static_cast<unsigned int>(nc);
  float best = -INFINITY;
  int bc = 0, cnt = 0;
  if (multi && !dedupe) {   // one candidate per (row, class) with conf > thr, general_yolo.py:191-193
    for (int c = lane; c < nc; c += 32) {
      const float raw = fe.cls_raw(b, h, c);
      if (!fe.cls_may_pass(raw, obj)) continue;
      const float conf = __fmul_rn(fe.cls(raw), obj);
      if (conf > p.conf_thres && class_ok(p.class_mask, c)) { f(cand_key(conf, ord0 + c)); ++cnt; }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, sft);
    return cnt;
  }
  // best class (first maximum); in dedupe mode only over the classes that are candidates themselves
  for (int c = lane; c < nc; c += 32) {
    const float raw = fe.cls_raw(b, h, c);
    if (!fe.cls_may_pass(raw, obj)) continue;   // cannot exceed the threshold, so it cannot be a maximum that matters
    const float conf = __fmul_rn(fe.cls(raw), obj);
    if (multi && !(conf > p.conf_thres && class_ok(p.class_mask, c))) continue;
    if (multi) ++cnt;
    if (conf > best) { best = conf; bc = c; }
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, sft);
    const int oc = __shfl_xor_sync(0xffffffffu, bc, sft);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, sft);
    if (ob > best || (ob == best && oc < bc)) { best = ob; bc = oc; }
  }
  if (!multi) {             // best class only, general_yolo.py:195-196
    const bool ok = best > p.conf_thres && class_ok(p.class_mask, bc);
    if (lane == 0 && ok) f(cand_key(best, ord0 + bc));
    return ok ? 1 : 0;
  }
  if (cnt >= 2) {
    int ok = 0;
    if (lane == 0) {
      const float4 bx = fe.box(b, h);
      const float w = __fsub_rn(bx.z, bx.x), hh = __fsub_rn(bx.w, bx.y);
      ok = (w > 0.0f && hh > 0.0f && __fmul_rn(w, hh) < 1e37f) ? 1 : 0;
    }
    if (!__shfl_sync(0xffffffffu, ok, 0)) {   // degenerate box: every class stays a candidate
      for (int c = lane; c < nc; c += 32) {
        const float raw = fe.cls_raw(b, h, c);
        if (!fe.cls_may_pass(raw, obj)) continue;
        const float conf = __fmul_rn(fe.cls(raw), obj);
        if (conf > p.conf_thres && class_ok(p.class_mask, c)) f(cand_key(conf, ord0 + c));
      }
      return cnt;
    }
  }
  if (lane == 0 && cnt >= 1) f(cand_key(best, ord0 + bc));
  return cnt;
}

struct NmsSmem {
  unsigned long long pass_row[kPassSmem];   // row handles of the objectness survivors
  float pass_obj[kPassSmem];
  unsigned long long key[kSmemCap];
  float4 box[kSmemCap];
  unsigned long long removed[kSmemCap / 64];
  unsigned long long diag[2][64];
  unsigned int hist[256];
  unsigned long long keep_bits, prefix;
  int n_pass, n_emit, n_real, n_keep[2], k_rem;   // n_keep[ch & 1]: boxes kept before chunk ch (double-buffered: written during the previous chunk)
};

template <class FE>
__global__ void __launch_bounds__(kNmsThreads, 1) box_nms_kernel(const FE fe, int B, YpNmsParams p, int cap, int cap_p2, NmsWs ws,
                                                                 float* __restrict__ out_boxes, int* __restrict__ out_count, int prescanned) {
  extern __shared__ __align__(16) unsigned char nms_smem_raw[];
  NmsSmem& sm = *reinterpret_cast<NmsSmem*>(nms_smem_raw);
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int A = static_cast<int>(fe.A), nc = fe.no - 5;
  const bool multi = p.multi_label && nc > 1;
  unsigned long long* g_pass_row = ws.pass_row + static_cast<int64_t>(b) * A;
  float* g_pass_obj = ws.pass_obj + static_cast<int64_t>(b) * A;
  if (tid == 0) { sm.n_pass = 0; sm.n_emit = 0; sm.n_real = 0; sm.n_keep[0] = 0; sm.n_keep[1] = 0; }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");   // the logits are the previous kernels' output
  __syncthreads();
  bool dedupe = multi && p.agnostic && p.iou_thres < 1.0f;   // see warp_row_candidates

  // Segments (Detect levels) whose candidates nms_prescan_kernel already listed while the rest of the network was still running:
  // take the list if it is complete, else scan everything here.
  int have = 0;
  if (prescanned) {
    const int pn = ws.pre_n[2 * b], pr = ws.pre_n[2 * b + 1];
    if (pn <= kSmemCap && pr <= p.max_nms) {
      have = prescanned;
      for (int i = tid; i < pn; i += kNmsThreads) sm.key[i] = ws.pre_key[static_cast<int64_t>(b) * kSmemCap + i];
      if (tid == 0) { sm.n_emit = pn; sm.n_real = pr; }
    }
    __syncthreads();
    if (tid == 0) { ws.pre_n[2 * b] = 0; ws.pre_n[2 * b + 1] = 0; }   // ready for the next frame
  }

  // ---- A: objectness scan (strict >, general_yolo.py:146); four independent loads in flight per thread
  auto scan_objectness = [&](int skip_mask) {
    for (int l = 0; l < fe.segments(); ++l) {
      if ((skip_mask >> l) & 1) continue;
      const int rows = fe.seg_rows(l);
      for (int i0 = tid; i0 < rows; i0 += 4 * kNmsThreads) {
        float raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + u * kNmsThreads; raw[u] = i < rows ? fe_obj_raw_at(fe, b, l, i) : -INFINITY; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!fe.obj_may_pass(raw[u])) continue;
          const float o = fe.obj(raw[u]);
          if (o > p.conf_thres) {
            const int slot = atomicAdd(&sm.n_pass, 1);
            const unsigned long long h = fe.handle(l, i0 + u * kNmsThreads);
            if (slot < kPassSmem) { sm.pass_row[slot] = h; sm.pass_obj[slot] = o; }
            else { g_pass_row[slot] = h; g_pass_obj[slot] = o; }
          }
        }
      }
    }
  };
  scan_objectness(have);
  __syncthreads();
  int n_pass = sm.n_pass;
  auto pass_row = [&](int i) { return i < kPassSmem ? sm.pass_row[i] : g_pass_row[i]; };
  auto pass_obj = [&](int i) { return i < kPassSmem ? sm.pass_obj[i] : g_pass_obj[i]; };

  // ---- B: candidates.  One warp per surviving row; f(key) once per candidate.
  auto for_each_candidate = [&](auto&& f) {
    for (int pi = warp; pi < n_pass; pi += kNmsThreads / 32) {
      const int cnt = warp_row_candidates(fe, b, pass_row(pi), pass_obj(pi), p, nc, multi, dedupe, lane, f);
      if (lane == 0 && cnt) atomicAdd(&sm.n_real, cnt);
    }
  };
  unsigned long long* g_key = ws.key + static_cast<int64_t>(b) * cap_p2;
  for_each_candidate([&](unsigned long long k) {
    const int slot = atomicAdd(&sm.n_emit, 1);
    if (slot < kSmemCap) sm.key[slot] = k;
  });
  __syncthreads();
  const int n_real = sm.n_real;                 // (row, class) pairs above the threshold = the reference's candidate count
  int n_cand = sm.n_emit;                       // candidates that enter the sort (fewer than n_real with dedupe)
  if (dedupe && n_real > p.max_nms) { dedupe = false; n_cand = n_real; }   // the top-max_nms cut is over ALL candidates
  int path = 0, n = n_cand;
  unsigned long long* key = sm.key;
  float4* box = sm.box;
  unsigned long long* removed = sm.removed;
  if (n_cand > kSmemCap) {
    key = g_key; box = ws.box + static_cast<int64_t>(b) * cap; removed = ws.removed + static_cast<int64_t>(b) * (cap_p2 / 64);
    __syncthreads();
    if (have) {                                 // the prescan list is being dropped: list the rows of its segments as well
      scan_objectness(~have);
      __syncthreads();
      n_pass = sm.n_pass;
    }
    if (tid == 0) sm.n_emit = 0;
    __syncthreads();
    if (n_cand <= cap) {
      path = 1;
      for_each_candidate([&](unsigned long long k) { key[atomicAdd(&sm.n_emit, 1)] = k; });
    } else if (cap >= p.max_nms) {
      // ---- C: radix select (8 bits per pass, most significant first) of the max_nms-th largest key; keys are unique
      path = 2;
      if (tid == 0) { sm.prefix = 0ull; sm.k_rem = p.max_nms; }
      for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) sm.hist[tid] = 0u;
        __syncthreads();
        const unsigned long long prefix = sm.prefix;
        for_each_candidate([&](unsigned long long k) {
          if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&sm.hist[(k >> shift) & 255ull], 1u);
        });
        __syncthreads();
        if (tid == 0) {
          int cum = 0, d = 255;
          for (; d > 0; --d) {
            if (cum + static_cast<int>(sm.hist[d]) >= sm.k_rem) break;
            cum += sm.hist[d];
          }
          sm.k_rem -= cum;
          sm.prefix = (prefix << 8) | static_cast<unsigned long long>(d);
        }
        __syncthreads();
      }
      const unsigned long long kstar = sm.prefix;
      for_each_candidate([&](unsigned long long k) { if (k >= kstar) key[atomicAdd(&sm.n_emit, 1)] = k; });
    } else {
      // the caller's buffer is smaller than max_nms and overflowed: report, the host grows `cap` (never silently truncated)
      if (tid == 0) {
        out_count[b] = -1 - n_cand;
        ws.stats[b * 4] = n_pass; ws.stats[b * 4 + 1] = n_real; ws.stats[b * 4 + 2] = 0; ws.stats[b * 4 + 3] = 3;
      }
      return;
    }
    __syncthreads();
    n = sm.n_emit;
  }

  // ---- D: bitonic sort, descending; padding keys are 0 (< every candidate key)
  int n_p2 = 64;
  while (n_p2 < n) n_p2 <<= 1;
  for (int i = n + tid; i < n_p2; i += kNmsThreads) key[i] = 0ull;
  __syncthreads();
  for (int k = 2; k <= n_p2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n_p2 >> 1); t += kNmsThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
        const int l = i | j;
        const unsigned long long a = key[i], c = key[l];
        const bool desc = (i & k) == 0;
        if (desc ? a < c : a > c) { key[i] = c; key[l] = a; }
      }
      __syncthreads();
    }
  }
  const int n_sorted = min(n, p.max_nms);

  // ---- E: NMS boxes of the sorted candidates (class offset unless agnostic, general_yolo.py:216-217)
  for (int i = tid; i < n_sorted; i += kNmsThreads) {
    const unsigned int ord = 0xffffffffu - static_cast<unsigned int>(key[i]);
    const unsigned int r = ord / nc, c = ord - r * nc;
    float4 bx = fe.box(b, fe.handle_of_row(r));
    if (!p.agnostic) {
      const float off = __fmul_rn(static_cast<float>(c), p.max_wh);
      bx = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
    }
    box[i] = bx;
  }
  for (int w = tid; w < (n_p2 >> 6); w += kNmsThreads) removed[w] = 0ull;
  __syncthreads();

  // ---- F: chunked greedy suppression, two barriers per chunk:
  //   (1) thread 0 resolves chunk ch sequentially from its 64 x 64 IoU bits (the reference's scan restricted to one chunk);
  //   (2) everybody: the chunk's kept boxes are written out and clear all later boxes; the IoU bits of chunk ch+1 are computed.
  float* out = out_boxes + static_cast<int64_t>(b) * p.max_det * 6;
  const int n_chunks = (n_sorted + 63) >> 6;
  // 64 x 64 IoU bits of a chunk: thread (i, g) tests box i against boxes 4g..4g+3 of the chunk; 16 lanes OR their nibbles.
  // diag[i] = the EARLIER boxes of the chunk (j < i) whose IoU with box i exceeds the threshold (IoU is symmetric).
  auto chunk_bits = [&](int ch) {
    const int base = ch << 6, lim = min(64, n_sorted - base);
    const int i = tid >> 4, g = tid & 15;
    unsigned long long bits = 0ull;
    if (i < lim) {
      const float4 me = box[base + i];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = 4 * g + e;
        if (j < i && iou_gt(box[base + j], me, p.iou_thres)) bits |= 1ull << j;
      }
    }
#pragma unroll
    for (int sft = 8; sft > 0; sft >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, sft);
    if (g == 0) sm.diag[ch & 1][i] = bits;
  };
  if (n_chunks) chunk_bits(0);
  __syncthreads();
  int ch = 0;
  for (; ch < n_chunks; ++ch) {
    const int base = ch << 6;
    const int lim = min(64, n_sorted - base);
    const int n_keep = sm.n_keep[ch & 1];
    if (n_keep >= p.max_det) break;   // uniform: slot ch & 1 was written during the previous chunk, before a barrier
    if (warp == 0) {
      // The reference's scan restricted to one chunk, as a fixed point over bit masks (lane l owns boxes l and l + 32):
      // a box is SUPPRESSED once an earlier overlapping box is kept, KEPT once every earlier overlapping box is suppressed.
      // Both are final and equal the sequential decisions; the lowest undecided box is always decidable, so it terminates.
      // (A single thread walking the kept boxes one by one cost ~80 ns per kept box: 10 us per frame.)
      const unsigned long long live = lim < 64 ? (1ull << lim) - 1ull : ~0ull;
      const unsigned long long p0 = sm.diag[ch & 1][lane], p1 = sm.diag[ch & 1][lane + 32];
      unsigned long long K = 0ull, U = live & ~removed[ch];
      while (U) {
        const bool u0 = (U >> lane) & 1ull, u1 = (U >> (lane + 32)) & 1ull;
        const bool s0 = u0 && (p0 & K), s1 = u1 && (p1 & K);
        const bool k0 = u0 && !s0 && !(p0 & U), k1 = u1 && !s1 && !(p1 & U);
        const unsigned long long nk = static_cast<unsigned long long>(__ballot_sync(0xffffffffu, k0)) |
                                      (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, k1)) << 32);
        const unsigned long long ns = static_cast<unsigned long long>(__ballot_sync(0xffffffffu, s0)) |
                                      (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, s1)) << 32);
        K |= nk;
        U &= ~(nk | ns);
      }
      // torchvision keeps all of them; general_yolo.py:219-220 then cuts the list at max_det
      int room = p.max_det - n_keep;
      unsigned long long kb = K;
      if (__popcll(K) > room) {   // drop the kept boxes beyond max_det (highest indices)
        kb = 0ull;
        unsigned long long m = K;
        for (; room > 0; --room) { const unsigned long long low = m & (~m + 1ull); kb |= low; m ^= low; }
      }
      if (lane == 0) { sm.keep_bits = kb; sm.n_keep[(ch + 1) & 1] = n_keep + __popcll(kb); }
    }
    __syncthreads();
    const unsigned long long kb = sm.keep_bits;
    if (tid < 64 && ((kb >> tid) & 1ull)) {   // kept rows, in order: (un-offset box, conf, cls)
      const int q = __popcll(kb & ((1ull << tid) - 1ull));
      const unsigned long long k = key[base + tid];
      const unsigned int ord = 0xffffffffu - static_cast<unsigned int>(k);
      const unsigned int r = ord / nc, c = ord - r * nc;
      const float4 bx = fe.box(b, fe.handle_of_row(r));
      float* o = out + static_cast<int64_t>(n_keep + q) * 6;
      o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = __uint_as_float(static_cast<unsigned int>(k >> 32)); o[5] = static_cast<float>(c);
    }
    if (ch + 1 < n_chunks) {
      if (kb) {
        for (int j = base + 64 + tid; j < n_sorted; j += kNmsThreads) {
          if ((removed[j >> 6] >> (j & 63)) & 1ull) continue;
          const float4 bj = box[j];
          bool hit = false;
          for (unsigned long long m = kb; m && !hit; m &= m - 1ull) hit = iou_gt(box[base + __ffsll(static_cast<long long>(m)) - 1], bj, p.iou_thres);
          if (hit) atomicOr(&removed[j >> 6], 1ull << (j & 63));
        }
      }
      chunk_bits(ch + 1);
    }
    __syncthreads();
  }
  if (tid == 0) {
    out_count[b] = sm.n_keep[ch & 1];   // after a break or the last chunk, slot ch & 1 holds the total
    ws.stats[b * 4] = n_pass; ws.stats[b * 4 + 1] = n_real; ws.stats[b * 4 + 2] = n_sorted; ws.stats[b * 4 + 3] = path;
  }
}

// Candidates of one segment (Detect level), listed ahead of the NMS kernel: grid-wide, one thread per row for the objectness test,
// then the warp evaluates the classes of its surviving rows one row at a time.  Levels 0 and 1 hold 95 % of the rows and are
// complete long before the last Detect convolution, so this runs under the rest of the network and the NMS kernel on the
// critical path starts from a finished list.  Emission order is irrelevant (the keys are sorted later); the counters are the only
// shared state and the NMS kernel zeroes them again.
template <class FE>
__global__ void __launch_bounds__(256) nms_prescan_kernel(const FE fe, int seg, YpNmsParams p, NmsWs ws) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int nc = fe.no - 5;
  const bool multi = p.multi_label && nc > 1;
  const bool dedupe = multi && p.agnostic && p.iou_thres < 1.0f;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int rows = fe.seg_rows(seg);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float o = 0.0f;
  bool pass = false;
  if (i < rows) {
    const float raw = fe_obj_raw_at(fe, b, seg, i);
    if (fe.obj_may_pass(raw)) { o = fe.obj(raw); pass = o > p.conf_thres; }
  }
  unsigned m = __ballot_sync(0xffffffffu, pass);
  int real = 0;
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const int ri = __shfl_sync(0xffffffffu, i, src);
    const float ro = __shfl_sync(0xffffffffu, o, src);
    real += warp_row_candidates(fe, b, fe.handle(seg, ri), ro, p, nc, multi, dedupe, lane, [&](unsigned long long k) {
      const int slot = atomicAdd(&ws.pre_n[2 * b], 1);
      if (slot < kSmemCap) ws.pre_key[static_cast<int64_t>(b) * kSmemCap + slot] = k;
    });
  }
  if (lane == 0 && real) atomicAdd(&ws.pre_n[2 * b + 1], real);
}

size_t align_up(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

int pow2_at_least(int v) { int p = 64; while (p < v) p <<= 1; return p; }

size_t carve(NmsWs* ws, char* base, int B, long long A, int cap) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
  const int cap_p2 = pow2_at_least(cap);
  ws->stats = reinterpret_cast<int*>(take(sizeof(int) * 4 * B));
  ws->pass_row = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * B * A));
  ws->pass_obj = reinterpret_cast<float*>(take(sizeof(float) * B * A));
  ws->key = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * B * cap_p2));
  ws->box = reinterpret_cast<float4*>(take(sizeof(float4) * B * cap));
  ws->removed = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * B * (cap_p2 / 64)));
  ws->pre_n = reinterpret_cast<int*>(take(sizeof(int) * 2 * B));
  ws->pre_key = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * B * kSmemCap));
  return off;
}

template <class FE>
int launch_box_nms(const FE& fe, int B, const YpNmsParams& p, int cap, const NmsWs& ws, float* out_boxes, int32_t* out_count, int prescanned,
                   cudaStream_t st) {
  auto kern = box_nms_kernel<FE>;
  static thread_local bool raised = false;
  if (!raised) { YP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(NmsSmem)))); raised = true; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B); cfg.blockDim = dim3(kNmsThreads); cfg.dynamicSmemBytes = sizeof(NmsSmem); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  YP_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, fe, B, p, cap, pow2_at_least(cap), ws, out_boxes, out_count, prescanned));
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

static int nms_check_params(const char* who, const YpNmsParams* p, int32_t cap) {
  YP_REQUIRE(cap > 0 && cap % 64 == 0, YP_ERR_SHAPE, "%s: cap=%d must be a positive multiple of 64", who, cap);
  YP_REQUIRE(p->conf_thres >= 0.f && p->conf_thres <= 1.f, YP_ERR_ARG, "Invalid Confidence threshold %g, valid values are between 0.0 and 1.0", p->conf_thres);
  YP_REQUIRE(p->iou_thres >= 0.f && p->iou_thres <= 1.f, YP_ERR_ARG, "Invalid IoU %g, valid values are between 0.0 and 1.0", p->iou_thres);
  YP_REQUIRE(p->max_det > 0 && p->max_nms > 0, YP_ERR_ARG, "%s: max_det/max_nms must be positive", who);
  return YP_OK;
}

extern "C" int yp_box_nms(const float* pred, int32_t B, int64_t A, int32_t no, const YpNmsParams* p, int32_t cap,
                          float* out_boxes, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
  YP_REQUIRE(pred && p && out_boxes && out_count && workspace, YP_ERR_ARG, "box_nms: null pointer");
  YP_REQUIRE(B > 0 && A > 0 && no >= 6, YP_ERR_SHAPE, "box_nms: B=%d A=%lld no=%d", B, (long long)A, no);
  YP_REQUIRE(A * (no - 5) < (1ll << 31), YP_ERR_SHAPE, "box_nms: A * nc = %lld exceeds 2^31", (long long)(A * (no - 5)));
  int rc = nms_check_params("box_nms", p, cap);
  if (rc != YP_OK) return rc;
  yp::NmsWs ws;
  const size_t need = yp::carve(&ws, static_cast<char*>(workspace), B, A, cap);
  YP_REQUIRE(workspace_bytes >= need, YP_ERR_CAPACITY, "box_nms: workspace %zu < %zu bytes", workspace_bytes, need);
  yp::PredRows fe;
  fe.pred = pred; fe.A = A; fe.no = no; fe.thr = p->conf_thres;
  return yp::launch_box_nms(fe, B, *p, cap, ws, out_boxes, out_count, 0, static_cast<cudaStream_t>(stream));
}

static int det_levels(const char* who, const float* const* logits3, const int32_t* ny3, const int32_t* nx3, const int32_t* ldc3, const float* stride3,
                      const float* anchors_px_host, int32_t B, int32_t na, int32_t no, const YpNmsParams* p, int32_t cap, yp::DetLevels* out) {
  YP_REQUIRE(logits3 && ny3 && nx3 && ldc3 && stride3 && anchors_px_host && p, YP_ERR_ARG, "%s: null pointer", who);
  YP_REQUIRE(B > 0 && na == 3 && no >= 6, YP_ERR_SHAPE, "%s: B=%d na=%d no=%d (na must be 3)", who, B, na, no);
  int rc = nms_check_params(who, p, cap);
  if (rc != YP_OK) return rc;
  yp::DetLevels& lv = *out;
  long long A = 0;
  for (int l = 0; l < 3; ++l) {
    YP_REQUIRE(logits3[l] && na * no <= ldc3[l], YP_ERR_SHAPE, "%s: level %d logits missing or ldc too small", who, l);
    YP_REQUIRE(ny3[l] > 0 && nx3[l] > 0 && ny3[l] < 16384 && nx3[l] < 16384, YP_ERR_SHAPE, "%s: level %d is %d x %d", who, l, ny3[l], nx3[l]);
    lv.logits[l] = logits3[l]; lv.ny[l] = ny3[l]; lv.nx[l] = nx3[l]; lv.ldc[l] = ldc3[l]; lv.stride[l] = stride3[l];
    for (int i = 0; i < 6; ++i) lv.anchor[l][i] = anchors_px_host[l * 6 + i];
    lv.row_off[l] = A;
    A += static_cast<long long>(na) * ny3[l] * nx3[l];
  }
  lv.na = na; lv.no = no; lv.A = A;
  {
    const double t = static_cast<double>(p->conf_thres) * (1.0 - 1e-4) - 1e-30;   // a threshold strictly below the real one
    lv.logit_thr = t <= 0.0 ? -INFINITY : static_cast<float>(log(t / (1.0 - t)) - 1e-4);
  }
  YP_REQUIRE(A * (no - 5) < (1ll << 31), YP_ERR_SHAPE, "%s: A * nc = %lld exceeds 2^31", who, (long long)(A * (no - 5)));
  return YP_OK;
}

extern "C" int yp_detect_prescan(const float* const* logits3, const int32_t* ny3, const int32_t* nx3, const int32_t* ldc3, const float* stride3,
                                 const float* anchors_px_host, int32_t B, int32_t na, int32_t no, const YpNmsParams* p, int32_t cap,
                                 int32_t level, void* workspace, size_t workspace_bytes, void* stream) {
  yp::DetLevels lv;
  int rc = det_levels("detect_prescan", logits3, ny3, nx3, ldc3, stride3, anchors_px_host, B, na, no, p, cap, &lv);
  if (rc != YP_OK) return rc;
  YP_REQUIRE(workspace && level >= 0 && level < 3, YP_ERR_ARG, "detect_prescan: level %d / workspace", level);
  yp::NmsWs ws;
  const size_t need = yp::carve(&ws, static_cast<char*>(workspace), B, lv.A, cap);
  YP_REQUIRE(workspace_bytes >= need, YP_ERR_CAPACITY, "detect_prescan: workspace %zu < %zu bytes", workspace_bytes, need);
  const int rows = 3 * ny3[level] * nx3[level];
  yp::nms_prescan_kernel<yp::DetLevels><<<dim3(yp::ceil_div(rows, 256), B), 256, 0, static_cast<cudaStream_t>(stream)>>>(lv, level, *p, ws);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_detect_nms(const float* const* logits3, const int32_t* ny3, const int32_t* nx3, const int32_t* ldc3, const float* stride3,
                             const float* anchors_px_host, int32_t B, int32_t na, int32_t no, const YpNmsParams* p, int32_t cap,
                             float* out_boxes, int32_t* out_count, void* workspace, size_t workspace_bytes, int32_t prescanned_levels,
                             void* stream) {
  YP_REQUIRE(out_boxes && out_count && workspace, YP_ERR_ARG, "detect_nms: null pointer");
  YP_REQUIRE(prescanned_levels >= 0 && prescanned_levels < 8, YP_ERR_ARG, "detect_nms: prescanned_levels=%d", prescanned_levels);
  yp::DetLevels lv;
  int rc = det_levels("detect_nms", logits3, ny3, nx3, ldc3, stride3, anchors_px_host, B, na, no, p, cap, &lv);
  if (rc != YP_OK) return rc;
  yp::NmsWs ws;
  const size_t need = yp::carve(&ws, static_cast<char*>(workspace), B, lv.A, cap);
  YP_REQUIRE(workspace_bytes >= need, YP_ERR_CAPACITY, "detect_nms: workspace %zu < %zu bytes", workspace_bytes, need);
  return yp::launch_box_nms(lv, B, *p, cap, ws, out_boxes, out_count, prescanned_levels, static_cast<cudaStream_t>(stream));
}
