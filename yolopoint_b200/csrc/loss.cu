// Training-side target / loss kernels: the keypoint-detector loss of the training step
// (src/train.py:220-232) with its label / mask preparation, forward AND backward in one pass over the logits:
//     labels2Dto3D (utils/utils.py:184-209: 8x8 pixel-unshuffle, dustbin channel, per-cell normalisation),
//     getMasks     (utils/utils.py:103-116: a cell is valid iff all 64 pixels are),
//     ComputeDetectorLoss (utils/loss_functions.py:600-619: BCE between softmax(semi) and the 65-channel labels, summed over
//                  the channels, masked mean over the cells).
// The reference runs ~25 ATen kernels over [B,65,Hc,Wc] / [B,1,H,W] tensors per call (twice per step) and autograd stores the
// softmax, the BCE terms and the labels for the backward pass; here one thread owns one cell (65 logits, 64 label pixels, 64 mask
// pixels in registers), emits the cell's loss term and the 65 gradients d loss / d semi.  HBM-bound: reads 65 + 64 + 64 floats
// per cell, writes 65.  Sums are reduced per block and added in a fixed order (bit-reproducible).
#include "common.cuh"

namespace yp {
namespace {

constexpr int kLossThreads = 128;

__device__ __forceinline__ float block_sum(float v, float* red) {
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

// mask3d[b, hc, wc] = prod of the 8x8 mask pixels; part[block] = sum of the block's cells
__global__ void __launch_bounds__(kLossThreads) det_mask_kernel(const float* __restrict__ mask2d, int B, int Hc, int Wc, float* __restrict__ mask3d,
                                                                 float* __restrict__ part) {
  __shared__ float red[kLossThreads / 32];
  const int64_t total = static_cast<int64_t>(B) * Hc * Wc;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  float m = 0.0f;
  if (idx < total) {
    const int wc = static_cast<int>(idx % Wc), hc = static_cast<int>((idx / Wc) % Hc), b = static_cast<int>(idx / (static_cast<int64_t>(Wc) * Hc));
    const int W = Wc * 8;
    const float* p = mask2d + (static_cast<int64_t>(b) * Hc * 8 + hc * 8) * W + wc * 8;
    m = 1.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p + static_cast<int64_t>(i) * W));
      const float4 c = __ldg(reinterpret_cast<const float4*>(p + static_cast<int64_t>(i) * W) + 1);
      m = m * a.x * a.y * a.z * a.w * c.x * c.y * c.z * c.w;
    }
    mask3d[idx] = m;
  }
  const float s = block_sum(m, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kLossThreads) det_loss_kernel(const float* __restrict__ semi, long long sB, long long sC, long long sH, long long sW,
                                                                 const float* __restrict__ labels2d, const float* __restrict__ mask3d,
                                                                 const float* __restrict__ mask_part, int n_mask_part, int B, int Hc, int Wc,
                                                                 float* __restrict__ dsemi, float* __restrict__ loss_part) {
  __shared__ float red[kLossThreads / 32];
  __shared__ float denom_s;
  if (threadIdx.x == 0) {   // every block adds the mask partial sums in the same order
    float t = 0.0f;
    for (int i = 0; i < n_mask_part; ++i) t += mask_part[i];
    denom_s = t + 1e-10f;
  }
  __syncthreads();
  const float inv_denom = 1.0f / denom_s;
  const int64_t total = static_cast<int64_t>(B) * Hc * Wc;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  float cell_loss = 0.0f;
  if (idx < total) {
    const int wc = static_cast<int>(idx % Wc), hc = static_cast<int>((idx / Wc) % Hc), b = static_cast<int>(idx / (static_cast<int64_t>(Wc) * Hc));
    const int W = Wc * 8;
    const float* lp = labels2d + (static_cast<int64_t>(b) * Hc * 8 + hc * 8) * W + wc * 8;
    float t[65];
    float lsum = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(lp + static_cast<int64_t>(i) * W));
      const float4 c = __ldg(reinterpret_cast<const float4*>(lp + static_cast<int64_t>(i) * W) + 1);
      t[8 * i] = a.x; t[8 * i + 1] = a.y; t[8 * i + 2] = a.z; t[8 * i + 3] = a.w;
      t[8 * i + 4] = c.x; t[8 * i + 5] = c.y; t[8 * i + 6] = c.z; t[8 * i + 7] = c.w;
    }
#pragma unroll
    for (int c = 0; c < 64; ++c) lsum += t[c];
    const float dust = (1.0f - lsum) < 1.0f ? 0.0f : 1.0f - lsum;   // utils.py:203-204
    t[64] = dust;
    const float tnorm = lsum + dust;
    const float* s = semi + b * sB + hc * sH + wc * sW;
    float p[65];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 65; ++c) { p[c] = __ldg(s + c * sC); mx = fmaxf(mx, p[c]); }
    float psum = 0.0f;
#pragma unroll
    for (int c = 0; c < 65; ++c) { p[c] = expf(p[c] - mx); psum += p[c]; }
    const float pinv = 1.0f / psum;
    const float m = mask3d[idx];
    // BCE on probabilities as F.binary_cross_entropy: logs clamped at -100; backward (p - t) / max(p (1 - p), 1e-12)
    float g[65];
    float dot = 0.0f;
#pragma unroll
    for (int c = 0; c < 65; ++c) {
      const float pc = p[c] * pinv, tc = t[c] / tnorm;
      p[c] = pc;
      cell_loss -= tc * fmaxf(logf(pc), -100.0f) + (1.0f - tc) * fmaxf(logf(1.0f - pc), -100.0f);
      g[c] = (pc - tc) / fmaxf(pc * (1.0f - pc), 1e-12f);
      dot += pc * g[c];
    }
    cell_loss *= m;
    const float scale = m * inv_denom;
    float* d = dsemi + b * sB + hc * sH + wc * sW;
#pragma unroll
    for (int c = 0; c < 65; ++c) d[c * sC] = scale * p[c] * (g[c] - dot);     // softmax backward
  }
  const float sum = block_sum(cell_loss, red);
  if (threadIdx.x == 0) loss_part[blockIdx.x] = sum;
}

__global__ void det_finalize_kernel(const float* __restrict__ loss_part, int n_loss, const float* __restrict__ mask_part, int n_mask, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float l = 0.0f, m = 0.0f;
    for (int i = 0; i < n_loss; ++i) l += loss_part[i];
    for (int i = 0; i < n_mask; ++i) m += mask_part[i];
    out[0] = l / (m + 1e-10f);
    out[1] = m;
  }
}

}  // namespace
}  // namespace yp

extern "C" size_t yp_detector_loss_workspace_bytes(int32_t B, int32_t Hc, int32_t Wc) {
  if (B <= 0 || Hc <= 0 || Wc <= 0) return 0;
  const int64_t cells = static_cast<int64_t>(B) * Hc * Wc;
  const int64_t blocks = (cells + yp::kLossThreads - 1) / yp::kLossThreads;
  return static_cast<size_t>(cells + 2 * blocks) * sizeof(float);
}

extern "C" int yp_detector_loss(const float* semi, int64_t sB, int64_t sC, int64_t sH, int64_t sW, const float* labels2d, const float* mask2d,
                                int32_t B, int32_t Hc, int32_t Wc, float* dsemi, float* out2, void* workspace, size_t workspace_bytes, void* stream) {
  YP_REQUIRE(semi && labels2d && mask2d && dsemi && out2 && workspace, YP_ERR_ARG, "detector_loss: null pointer");
  YP_REQUIRE(B > 0 && Hc > 0 && Wc > 0, YP_ERR_SHAPE, "detector_loss: bad shape");
  YP_REQUIRE(yp::aligned16(labels2d) && yp::aligned16(mask2d) && (Wc * 8) % 4 == 0, YP_ERR_ALIGN, "detector_loss: labels / mask must be 16-byte aligned");
  const size_t need = yp_detector_loss_workspace_bytes(B, Hc, Wc);
  YP_REQUIRE(workspace_bytes >= need, YP_ERR_CAPACITY, "detector_loss: workspace %zu < %zu bytes", workspace_bytes, need);
  const int64_t cells = static_cast<int64_t>(B) * Hc * Wc;
  const int blocks = static_cast<int>((cells + yp::kLossThreads - 1) / yp::kLossThreads);
  float* mask3d = static_cast<float*>(workspace);
  float* mask_part = mask3d + cells;
  float* loss_part = mask_part + blocks;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  yp::det_mask_kernel<<<blocks, yp::kLossThreads, 0, st>>>(mask2d, B, Hc, Wc, mask3d, mask_part);
  yp::det_loss_kernel<<<blocks, yp::kLossThreads, 0, st>>>(semi, sB, sC, sH, sW, labels2d, mask3d, mask_part, blocks, B, Hc, Wc, dsemi, loss_part);
  yp::det_finalize_kernel<<<1, 32, 0, st>>>(loss_part, blocks, mask_part, blocks, out2);
  YP_LAUNCH_OK();
  return YP_OK;
}
