// Training-mode BatchNorm2d + SiLU of a Conv block (models/common.py:22-34 of the reference: act(bn(conv(x))) with batch statistics)
// on bf16 NHWC activations: forward (statistics, normalise + activate) and backward (reductions, input gradient).
//
// All four passes are HBM-bound streaming kernels over a [P, C] matrix (P = B*H*W pixels, C channels contiguous); a thread owns one
// group of 8 channels (one 16-byte vector) and walks the pixels with a stride, so every access is a coalesced 16-byte load/store and
// the per-channel partial sums live in registers.  Per-block partials are combined in shared memory and added to fp32 accumulators
// in global memory (2*C floats per pass).
//   forward : read y (2 B/elem) for the statistics; read y, write out (4 B/elem) for normalise + SiLU
//   backward: read dout, y (4 B/elem) for sum(dz), sum(dz * xhat); read dout, y, write dy (6 B/elem)
#include "common.cuh"

namespace yp {
namespace {

constexpr int kBnThreads = 256;
constexpr int kBnUnroll = 4;   // pixel rows (independent 16-byte loads per operand) a thread keeps in flight

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

// Thread layout shared by the passes: G = C/8 channel groups; a block covers `gpb` consecutive groups x `rpb` pixel rows per
// iteration (gpb * rpb = 256); blockIdx.y walks the group tiles, blockIdx.x the pixel chunks.
struct Layout {
  int gpb, rpb;
};
__host__ __device__ inline Layout make_layout(int C) {
  const int G = C / 8;
  int gpb = 1;
  while (gpb * 2 <= G && gpb * 2 <= kBnThreads && G % (gpb * 2) == 0) gpb *= 2;   // largest power of two dividing G (<= 256)
  return Layout{gpb, kBnThreads / gpb};
}

// two per-channel sums over the pixels: MODE 0 -> (sum y, sum y^2); MODE 1 -> (sum dz, sum dz * xhat)
template <int MODE>
__global__ void __launch_bounds__(kBnThreads) bn_reduce_kernel(const uint4* __restrict__ y, const uint4* __restrict__ dout, long long P, int C,
                                                               const float* __restrict__ save, int act, float* __restrict__ acc, int rows_per_block) {
  __shared__ float red[2][kBnThreads][8 + 1];
  const Layout L = make_layout(C);
  const int G = C / 8;
  const int gl = threadIdx.x % L.gpb, rl = threadIdx.x / L.gpb;
  const int g = blockIdx.y * L.gpb + gl;
  const long long p0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long p1 = p0 + rows_per_block < P ? p0 + rows_per_block : P;
  float s0[8], s1[8], mean[8], rstd[8], scale[8], shift[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s0[e] = 0.f; s1[e] = 0.f; }
  if (MODE == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mean[e] = save[g * 8 + e]; rstd[e] = save[C + g * 8 + e]; scale[e] = save[2 * C + g * 8 + e]; shift[e] = save[3 * C + g * 8 + e];
    }
  }
  auto row = [&](const uint4& yu, const uint4& du) {
    float v[8];
    unpack8(yu, v);
    if (MODE == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { s0[e] += v[e]; s1[e] = fmaf(v[e], v[e], s1[e]); }
    } else {
      float d[8];
      unpack8(du, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float dz = d[e];
        if (act) {
          const float z = fmaf(v[e], scale[e], shift[e]);
          const float s = __fdividef(1.0f, 1.0f + __expf(-z));
          dz *= s * fmaf(z, 1.0f - s, 1.0f);
        }
        s0[e] += dz;
        s1[e] = fmaf(dz, (v[e] - mean[e]) * rstd[e], s1[e]);
      }
    }
  };
  // kBnUnroll independent 16-byte loads per operand in flight per thread: one load per iteration left the passes latency-bound
  // (51 % / 33 % of the HBM peak for the forward / backward pair, tools/postproc_roofline.py)
  long long p = p0 + rl;
  const long long step = L.rpb;
  for (; p + (kBnUnroll - 1) * step < p1; p += kBnUnroll * step) {
    uint4 yu[kBnUnroll], du[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      yu[u] = __ldg(y + (p + u * step) * G + g);
      du[u] = MODE == 1 ? __ldg(dout + (p + u * step) * G + g) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) row(yu[u], du[u]);
  }
  for (; p < p1; p += step) row(__ldg(y + p * G + g), MODE == 1 ? __ldg(dout + p * G + g) : make_uint4(0u, 0u, 0u, 0u));
#pragma unroll
  for (int e = 0; e < 8; ++e) { red[0][threadIdx.x][e] = s0[e]; red[1][threadIdx.x][e] = s1[e]; }
  __syncthreads();
  // thread t < gpb * 8 * 2 sums one (which, group, channel) column over the rpb row-threads
  for (int t = threadIdx.x; t < L.gpb * 16; t += kBnThreads) {
    const int which = t / (L.gpb * 8), rem = t % (L.gpb * 8), gg = rem / 8, e = rem % 8;
    float s = 0.f;
    for (int r = 0; r < L.rpb; ++r) s += red[which][r * L.gpb + gg][e];
    atomicAdd(acc + which * C + (blockIdx.y * L.gpb + gg) * 8 + e, s);
  }
}

// per channel: batch statistics -> (mean, rstd, scale, shift).  Evaluated by every thread of the normalise pass for its own 8 channels
// (a few fp64 operations per thread; a separate finalize launch per BatchNorm cost more in launch gaps than the arithmetic); the
// threads of the first pixel chunk also publish the four vectors for the backward pass and update the running statistics
// (momentum, unbiased variance).
struct BnStats {
  float mean, rstd, scale, shift;
  double var;
};
__device__ __forceinline__ BnStats bn_channel_stats(const float* __restrict__ acc, long long P, int C, int c, float gamma, float beta, float eps) {
  const double n = static_cast<double>(P);
  const double m = acc[c] / n;
  double var = acc[C + c] / n - m * m;
  if (var < 0.0) var = 0.0;
  BnStats r;
  r.var = var;
  r.mean = static_cast<float>(m);
  r.rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  r.scale = gamma * r.rstd;
  r.shift = beta - r.mean * r.scale;
  return r;
}

__global__ void __launch_bounds__(kBnThreads) bn_apply_fwd_kernel(const uint4* __restrict__ y, long long P, int C, const float* __restrict__ acc,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
                                                                  float* __restrict__ save, int act, uint4* __restrict__ out, int rows_per_block) {
  const Layout L = make_layout(C);
  const int G = C / 8;
  const int gl = threadIdx.x % L.gpb, rl = threadIdx.x / L.gpb;
  const int g = blockIdx.y * L.gpb + gl;
  const long long p0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long p1 = p0 + rows_per_block < P ? p0 + rows_per_block : P;
  float scale[8], shift[8];
  const bool publish = blockIdx.x == 0 && rl == 0;      // one thread per channel group
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = g * 8 + e;
    const BnStats st = bn_channel_stats(acc, P, C, c, gamma[c], beta[c], eps);
    scale[e] = st.scale; shift[e] = st.shift;
    if (publish) {
      save[c] = st.mean; save[C + c] = st.rstd; save[2 * C + c] = st.scale; save[3 * C + c] = st.shift;
      if (running_mean) {
        const double n = static_cast<double>(P);
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * st.mean;
        const double unbiased = P > 1 ? st.var * n / (n - 1.0) : st.var;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
      }
    }
  }
  auto row = [&](const uint4& yu, long long p) {
    float v[8];
    unpack8(yu, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float z = fmaf(v[e], scale[e], shift[e]);
      v[e] = act ? __fdividef(z, 1.0f + __expf(-z)) : z;
    }
    out[p * G + g] = pack8(v);
  };
  long long p = p0 + rl;
  const long long step = L.rpb;
  for (; p + (kBnUnroll - 1) * step < p1; p += kBnUnroll * step) {
    uint4 yu[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) yu[u] = __ldg(y + (p + u * step) * G + g);
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) row(yu[u], p + u * step);
  }
  for (; p < p1; p += step) row(__ldg(y + p * G + g), p);
}

__global__ void __launch_bounds__(kBnThreads) bn_apply_bwd_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ y, long long P, int C,
                                                                  const float* __restrict__ gamma, const float* __restrict__ save, const float* __restrict__ acc,
                                                                  int act, uint4* __restrict__ dy, int rows_per_block) {
  const Layout L = make_layout(C);
  const int G = C / 8;
  const int gl = threadIdx.x % L.gpb, rl = threadIdx.x / L.gpb;
  const int g = blockIdx.y * L.gpb + gl;
  const long long p0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long p1 = p0 + rows_per_block < P ? p0 + rows_per_block : P;
  const float inv_n = 1.0f / static_cast<float>(P);
  float mean[8], rstd[8], scale[8], shift[8], k0[8], k1[8], k2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = g * 8 + e;
    mean[e] = save[c]; rstd[e] = save[C + c]; scale[e] = save[2 * C + c]; shift[e] = save[3 * C + c];
    k0[e] = gamma[c] * rstd[e];              // dy = k0 * (dz - k1 - xhat * k2)
    k1[e] = acc[c] * inv_n;
    k2[e] = acc[C + c] * inv_n;
  }
  auto row = [&](const uint4& yu, const uint4& du, long long p) {
    float v[8], d[8];
    unpack8(yu, v);
    unpack8(du, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float dz = d[e];
      if (act) {
        const float z = fmaf(v[e], scale[e], shift[e]);
        const float s = __fdividef(1.0f, 1.0f + __expf(-z));
        dz *= s * fmaf(z, 1.0f - s, 1.0f);
      }
      const float xhat = (v[e] - mean[e]) * rstd[e];
      v[e] = k0[e] * (dz - k1[e] - xhat * k2[e]);
    }
    dy[p * G + g] = pack8(v);
  };
  long long p = p0 + rl;
  const long long step = L.rpb;
  for (; p + (kBnUnroll - 1) * step < p1; p += kBnUnroll * step) {
    uint4 yu[kBnUnroll], du[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) { yu[u] = __ldg(y + (p + u * step) * G + g); du[u] = __ldg(dout + (p + u * step) * G + g); }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) row(yu[u], du[u], p + u * step);
  }
  for (; p < p1; p += step) row(__ldg(y + p * G + g), __ldg(dout + p * G + g), p);
}

// grid of the streaming passes: blockIdx.y walks the channel-group tiles, blockIdx.x contiguous pixel ranges (~`waves` blocks per SM)
dim3 pass_grid(long long P, int C, int waves, int* rows_per_block) {
  const Layout L = make_layout(C);
  const int gy = (C / 8) / L.gpb;
  long long bx = (static_cast<long long>(sm_count()) * waves + gy - 1) / gy;
  const long long max_bx = (P + L.rpb - 1) / L.rpb;
  if (bx > max_bx) bx = max_bx;
  if (bx < 1) bx = 1;
  const int rows = static_cast<int>((P + bx - 1) / bx);
  *rows_per_block = rows;
  return dim3(static_cast<unsigned>((P + rows - 1) / rows), gy);
}

int launch_reduce(int mode, const void* y, const void* dout, long long P, int C, const float* save, int act, float* acc, cudaStream_t st) {
  int rows = 0;
  const dim3 grid = pass_grid(P, C, 4, &rows);
  if (mode == 0)
    bn_reduce_kernel<0><<<grid, kBnThreads, 0, st>>>(static_cast<const uint4*>(y), nullptr, P, C, nullptr, act, acc, rows);
  else
    bn_reduce_kernel<1><<<grid, kBnThreads, 0, st>>>(static_cast<const uint4*>(y), static_cast<const uint4*>(dout), P, C, save, act, acc, rows);
  YP_LAUNCH_OK();
  return YP_OK;
}

}  // namespace
}  // namespace yp

extern "C" int yp_bn_act_fwd(const void* y, int64_t P, int32_t C, const float* gamma, const float* beta, float* running_mean, float* running_var,
                             float momentum, float eps, int32_t act, void* out, float* save, float* acc, void* stream) {
  YP_REQUIRE(y && gamma && beta && out && save && acc, YP_ERR_ARG, "bn_act_fwd: null pointer");
  YP_REQUIRE(P > 0 && C > 0 && C % 8 == 0, YP_ERR_SHAPE, "bn_act_fwd: P=%lld C=%d (C must be a multiple of 8)", static_cast<long long>(P), C);
  YP_REQUIRE(yp::aligned16(y) && yp::aligned16(out), YP_ERR_ALIGN, "bn_act_fwd: activations not 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  YP_CUDA_OK(cudaMemsetAsync(acc, 0, 2 * sizeof(float) * C, st));
  int rc = yp::launch_reduce(0, y, nullptr, P, C, nullptr, act, acc, st);
  if (rc != YP_OK) return rc;
  int rows = 0;
  const dim3 grid = yp::pass_grid(P, C, 8, &rows);
  yp::bn_apply_fwd_kernel<<<grid, yp::kBnThreads, 0, st>>>(static_cast<const uint4*>(y), P, C, acc, gamma, beta, running_mean, running_var, momentum, eps, save,
                                                        act, static_cast<uint4*>(out), rows);
  YP_LAUNCH_OK();
  return YP_OK;
}

extern "C" int yp_bn_act_bwd(const void* dout, const void* y, int64_t P, int32_t C, const float* gamma, const float* save, int32_t act, void* dy,
                             float* dgamma_dbeta, void* stream) {
  YP_REQUIRE(dout && y && gamma && save && dy && dgamma_dbeta, YP_ERR_ARG, "bn_act_bwd: null pointer");
  YP_REQUIRE(P > 0 && C > 0 && C % 8 == 0, YP_ERR_SHAPE, "bn_act_bwd: P=%lld C=%d (C must be a multiple of 8)", static_cast<long long>(P), C);
  YP_REQUIRE(yp::aligned16(y) && yp::aligned16(dout) && yp::aligned16(dy), YP_ERR_ALIGN, "bn_act_bwd: activations not 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // dgamma_dbeta = [sum dz (= dbeta) | sum dz * xhat (= dgamma)], also the two reductions the input gradient needs
  YP_CUDA_OK(cudaMemsetAsync(dgamma_dbeta, 0, 2 * sizeof(float) * C, st));
  int rc = yp::launch_reduce(1, y, dout, P, C, save, act, dgamma_dbeta, st);
  if (rc != YP_OK) return rc;
  int rows = 0;
  const dim3 grid = yp::pass_grid(P, C, 8, &rows);
  yp::bn_apply_bwd_kernel<<<grid, yp::kBnThreads, 0, st>>>(static_cast<const uint4*>(dout), static_cast<const uint4*>(y), P, C, gamma, save, dgamma_dbeta, act,
                                                        static_cast<uint4*>(dy), rows);
  YP_LAUNCH_OK();
  return YP_OK;
}
