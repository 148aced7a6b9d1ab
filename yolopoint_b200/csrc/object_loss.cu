// Training-side target / loss kernels, second half: the object loss of the training step
// (ComputeObjectLoss, src/utils/loss_functions.py:90-234: CIoU box loss, BCE objectness against the detached CIoU, BCE classes)
// over ALL Detect levels, forward AND gradient, in four small launches:
//
//   claim     one thread per assignment candidate (5 offsets x anchors x targets per level; the assignment rule of build_targets,
//             :218-234, is evaluated here from the label list -- or read from a host-built plan): the LAST valid candidate of a cell in candidate order owns the cell's objectness target
//             (the rule of `tobj[b, a, gj, gi] = iou`, :196, under duplicate indices) -- atomicMax on the candidate index;
//             counts the valid candidates of each level
//   candidate gather of the candidate's prediction row, box decode, CIoU and its gradient (box_loss_math.cuh), class BCE and its
//             gradient, objectness target of the cells it owns; gradients are scattered into d pred with atomicAdd
//   cells     one thread per (image, anchor, cell): objectness BCE against the target, its gradient into column 4 of d pred
//   finalize  fixed-order sums of the per-block partial losses -> (loss, box, obj, cls)
//
// The PyTorch form of the same loss (losses.ComputeObjectLoss.__call__, the checker in tests/test_gpu_train.py) runs ~45 ATen
// kernels per level forward and as many backward.  Latency-bound at training sizes (10^3 candidates, 2 x 10^5 cells): the only
// bandwidth term is the d pred buffer (cells x no floats, zero-filled then written in column 4).
// Loss VALUES are bit-reproducible (fixed-order partial sums); gradients of rows claimed by several candidates are summed with
// floating-point atomics (order-dependent in the last bit, as PyTorch's own index_put(accumulate) backward is).
#include "box_loss_math.cuh"
#include "common.cuh"

namespace yp {
namespace {

constexpr int kObjThreads = 128;
constexpr int kMaxLevels = YP_OBJ_LOSS_MAX_LEVELS;

struct ObjLevels {
  YpObjLossLevel lv[kMaxLevels];
  int64_t cand_off[kMaxLevels + 1];   // prefix sums of E
  int64_t cell_off[kMaxLevels + 1];   // prefix sums of cells
  int nl, no, nc;
  float cp, cn, cls_pw, obj_pw, gr, w_box, w_obj, w_cls, eps, anchor_t;
};

__device__ __forceinline__ float block_sum_f(float v, float* red) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
  __syncthreads();
  return t;   // valid in thread 0
}

// level of a flat index given prefix sums; a block never straddles levels (offsets are rounded up to the block size by the host)
__device__ __forceinline__ int level_of(const int64_t* off, int nl, int64_t idx) {
  int l = 0;
  while (l + 1 < nl && idx >= off[l + 1]) ++l;
  return l;
}

// candidate e of a level: from the host-built plan arrays, or (lv.targets) assigned here from the label list
struct Cand {
  bool valid;
  int64_t cell;
  float tbox[4];
  float aw, ah;
  int cls;
};

__device__ __forceinline__ Cand load_candidate(const YpObjLossLevel& lv, int64_t e, float anchor_t) {
  Cand c;
  c.valid = false;
  if (e >= lv.E) return c;
  if (lv.targets != nullptr) {
    const int t = static_cast<int>(e % lv.nt), a = static_cast<int>((e / lv.nt) % lv.na), o = static_cast<int>(e / (static_cast<int64_t>(lv.nt) * lv.na));
    c.aw = lv.anchors[2 * a];
    c.ah = lv.anchors[2 * a + 1];
    const CandPlan p = plan_candidate(lv.targets + 6 * static_cast<int64_t>(t), c.aw, c.ah, lv.nx, lv.ny, o, anchor_t);
    c.valid = p.valid && p.img >= 0 && p.img < lv.nb;
    c.cell = ((static_cast<int64_t>(p.img) * lv.na + a) * lv.ny + p.gj) * lv.nx + p.gi;
    c.cls = p.cls;
#pragma unroll
    for (int k = 0; k < 4; ++k) c.tbox[k] = p.tbox[k];
    return c;
  }
  if (lv.valid[e] == 0) return c;
  c.cell = lv.cell[e];
  c.valid = c.cell >= 0 && c.cell < lv.cells;
  c.aw = lv.anchor[2 * e];
  c.ah = lv.anchor[2 * e + 1];
  c.cls = static_cast<int>(lv.cls[e]);
#pragma unroll
  for (int k = 0; k < 4; ++k) c.tbox[k] = lv.tbox[4 * e + k];
  return c;
}

__global__ void __launch_bounds__(kObjThreads) obj_claim_kernel(const ObjLevels L, int* __restrict__ owner, int* __restrict__ n_valid) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int l = level_of(L.cand_off, L.nl, idx);
  const int64_t e = idx - L.cand_off[l];
  const YpObjLossLevel& lv = L.lv[l];
  const Cand c = load_candidate(lv, e, L.anchor_t);
  const bool ok = c.valid;
  if (ok) atomicMax(owner + L.cell_off[l] + c.cell, static_cast<int>(e));
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_valid + l, __popc(m));
}

__global__ void __launch_bounds__(kObjThreads) obj_candidate_kernel(const ObjLevels L, const int* __restrict__ owner, const int* __restrict__ n_valid,
                                                                     float* __restrict__ tobj, float* __restrict__ part_box, float* __restrict__ part_cls) {
  __shared__ float red[kObjThreads / 32];
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int l = level_of(L.cand_off, L.nl, idx);
  const int64_t e = idx - L.cand_off[l];
  const YpObjLossLevel& lv = L.lv[l];
  float box_term = 0.0f, cls_term = 0.0f;
  const Cand cd = load_candidate(lv, e, L.anchor_t);
  if (cd.valid) {
    const int n = n_valid[l];
    const int64_t cell = cd.cell;
    const float* q = lv.pred + cell * L.no;
    float* dq = lv.dpred + cell * L.no;
    float g[4];
    const float ciou = candidate_ciou(q, cd.aw, cd.ah, cd.tbox, L.eps, g);
    box_term = 1.0f - ciou;
    const float sb = -L.w_box / static_cast<float>(n);   // d (w_box * mean(1 - ciou)) / d ciou
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(dq + k, sb * g[k]);
    if (owner[L.cell_off[l] + cell] == static_cast<int>(e)) {
      const float score = fmaxf(ciou, 0.0f);
      tobj[L.cell_off[l] + cell] = (1.0f - L.gr) + L.gr * score;
    }
    if (L.nc > 1) {
      const int cls = cd.cls;
      const float sc = L.w_cls / (static_cast<float>(n) * static_cast<float>(L.nc));
      for (int c = 0; c < L.nc; ++c) {
        float dx;
        cls_term += bce_logits(q[5 + c], c == cls ? L.cp : L.cn, L.cls_pw, &dx);
        atomicAdd(dq + 5 + c, sc * dx);
      }
    }
  }
  const float sb = block_sum_f(box_term, red);
  const float sc = block_sum_f(cls_term, red);
  if (threadIdx.x == 0) {
    part_box[blockIdx.x] = sb;
    part_cls[blockIdx.x] = sc;
  }
}

__global__ void __launch_bounds__(kObjThreads) obj_cells_kernel(const ObjLevels L, const float* __restrict__ tobj, float* __restrict__ part_obj) {
  __shared__ float red[kObjThreads / 32];
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int l = level_of(L.cell_off, L.nl, idx);
  const int64_t c = idx - L.cell_off[l];
  const YpObjLossLevel& lv = L.lv[l];
  float term = 0.0f;
  if (c < lv.cells) {
    float dx;
    term = bce_logits(__ldg(lv.pred + c * L.no + 4), tobj[idx], L.obj_pw, &dx);
    lv.dpred[c * L.no + 4] = dx * (L.w_obj * lv.balance / static_cast<float>(lv.cells));
  }
  const float s = block_sum_f(term, red);
  if (threadIdx.x == 0) part_obj[blockIdx.x] = s;
}

// one warp: every lane adds its strided share of the partials in index order, then a shuffle tree (a fixed order)
__device__ __forceinline__ float warp_ordered_sum(const float* __restrict__ part, int64_t lo, int64_t hi) {
  float v = 0.0f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 32) v += part[i];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

__global__ void __launch_bounds__(32) obj_finalize_kernel(const ObjLevels L, const int* __restrict__ n_valid, const float* __restrict__ part_box,
                                                          const float* __restrict__ part_cls, const float* __restrict__ part_obj, float* __restrict__ out4) {
  float lbox = 0.0f, lcls = 0.0f, lobj = 0.0f;
  for (int l = 0; l < L.nl; ++l) {
    const int n = n_valid[l];
    const float b = warp_ordered_sum(part_box, L.cand_off[l] / kObjThreads, L.cand_off[l + 1] / kObjThreads);
    const float c = warp_ordered_sum(part_cls, L.cand_off[l] / kObjThreads, L.cand_off[l + 1] / kObjThreads);
    const float o = warp_ordered_sum(part_obj, L.cell_off[l] / kObjThreads, L.cell_off[l + 1] / kObjThreads);
    if (n > 0) {
      lbox += b / static_cast<float>(n);
      if (L.nc > 1) lcls += c / (static_cast<float>(n) * static_cast<float>(L.nc));
    }
    lobj += o / static_cast<float>(L.lv[l].cells) * L.lv[l].balance;
  }
  lbox *= L.w_box; lobj *= L.w_obj; lcls *= L.w_cls;
  if (threadIdx.x == 0) {
    out4[0] = lbox + lobj + lcls;
    out4[1] = lbox; out4[2] = lobj; out4[3] = lcls;
  }
}

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

}  // namespace
}  // namespace yp

extern "C" size_t yp_object_loss_workspace_bytes(const YpObjLossLevel* levels, int32_t nl) {
  if (!levels || nl <= 0 || nl > YP_OBJ_LOSS_MAX_LEVELS) return 0;
  int64_t cand = 0, cells = 0;
  for (int l = 0; l < nl; ++l) {
    cand += yp::round_up(levels[l].E, yp::kObjThreads);
    cells += yp::round_up(levels[l].cells, yp::kObjThreads);
  }
  // owner int32[cells] | tobj f32[cells] | n_valid int32[8] | part_box, part_cls f32[cand blocks] | part_obj f32[cell blocks]
  return static_cast<size_t>(2 * cells + 8 + 2 * (cand / yp::kObjThreads) + cells / yp::kObjThreads) * 4;
}

extern "C" int yp_object_loss(const YpObjLossLevel* levels, int32_t nl, int32_t no, int32_t nc, const YpObjLossParams* hp, float* out4,
                              void* workspace, size_t workspace_bytes, void* stream) {
  YP_REQUIRE(levels && hp && out4 && workspace, YP_ERR_ARG, "object_loss: null pointer");
  YP_REQUIRE(nl > 0 && nl <= YP_OBJ_LOSS_MAX_LEVELS && nc >= 1 && no == nc + 5, YP_ERR_SHAPE, "object_loss: nl=%d no=%d nc=%d", nl, no, nc);
  yp::ObjLevels L;
  memset(&L, 0, sizeof(L));
  L.nl = nl; L.no = no; L.nc = nc;
  L.cp = hp->cp; L.cn = hp->cn; L.cls_pw = hp->cls_pw; L.obj_pw = hp->obj_pw; L.gr = hp->gr;
  L.w_box = hp->w_box; L.w_obj = hp->w_obj; L.w_cls = hp->w_cls; L.eps = hp->eps; L.anchor_t = hp->anchor_t;
  for (int l = 0; l < nl; ++l) {
    const YpObjLossLevel& lv = levels[l];
    YP_REQUIRE(lv.pred && lv.dpred && lv.cells > 0 && lv.cells < (int64_t(1) << 31) && lv.E >= 0, YP_ERR_ARG, "object_loss: level %d: bad prediction buffers", l);
    if (lv.targets != nullptr)
      YP_REQUIRE(lv.na >= 1 && lv.na <= YP_OBJ_LOSS_MAX_ANCHORS && lv.nt >= 0 && lv.nx > 0 && lv.ny > 0 && lv.nb > 0 &&
                     lv.cells == static_cast<int64_t>(lv.nb) * lv.na * lv.ny * lv.nx && static_cast<int64_t>(lv.E) == 5LL * lv.na * lv.nt,
                 YP_ERR_SHAPE, "object_loss: level %d: nb=%d na=%d ny=%d nx=%d nt=%d do not match cells=%lld E=%d", l, lv.nb, lv.na, lv.ny, lv.nx, lv.nt,
                 static_cast<long long>(lv.cells), lv.E);
    else
      YP_REQUIRE(lv.E == 0 || (lv.valid && lv.cell && lv.tbox && lv.anchor && lv.cls), YP_ERR_ARG, "object_loss: level %d: null target plan", l);
    L.lv[l] = lv;
    L.cand_off[l + 1] = L.cand_off[l] + yp::round_up(lv.E, yp::kObjThreads);
    L.cell_off[l + 1] = L.cell_off[l] + yp::round_up(lv.cells, yp::kObjThreads);
  }
  for (int l = nl; l < yp::kMaxLevels; ++l) { L.cand_off[l + 1] = L.cand_off[nl]; L.cell_off[l + 1] = L.cell_off[nl]; }
  const size_t need = yp_object_loss_workspace_bytes(levels, nl);
  YP_REQUIRE(workspace_bytes >= need, YP_ERR_CAPACITY, "object_loss: workspace %zu < %zu bytes", workspace_bytes, need);
  const int64_t cells = L.cell_off[nl], cand = L.cand_off[nl];
  const int cand_blocks = static_cast<int>(cand / yp::kObjThreads), cell_blocks = static_cast<int>(cells / yp::kObjThreads);
  int* owner = static_cast<int*>(workspace);
  float* tobj = reinterpret_cast<float*>(owner + cells);
  int* n_valid = reinterpret_cast<int*>(tobj + cells);
  float* part_box = reinterpret_cast<float*>(n_valid + 8);
  float* part_cls = part_box + cand_blocks;
  float* part_obj = part_cls + cand_blocks;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  YP_CUDA_OK(cudaMemsetAsync(owner, 0xFF, static_cast<size_t>(cells) * 4, st));                 // -1: no owner
  YP_CUDA_OK(cudaMemsetAsync(tobj, 0, static_cast<size_t>(cells) * 4 + 8 * 4, st));              // targets and the level counters
  for (int l = 0; l < nl; ++l)
    YP_CUDA_OK(cudaMemsetAsync(levels[l].dpred, 0, static_cast<size_t>(levels[l].cells) * no * sizeof(float), st));
  if (cand_blocks > 0) {
    yp::obj_claim_kernel<<<cand_blocks, yp::kObjThreads, 0, st>>>(L, owner, n_valid);
    yp::obj_candidate_kernel<<<cand_blocks, yp::kObjThreads, 0, st>>>(L, owner, n_valid, tobj, part_box, part_cls);
  }
  yp::obj_cells_kernel<<<cell_blocks, yp::kObjThreads, 0, st>>>(L, tobj, part_obj);
  yp::obj_finalize_kernel<<<1, 32, 0, st>>>(L, n_valid, part_box, part_cls, part_obj, out4);
  YP_LAUNCH_OK();
  return YP_OK;
}
