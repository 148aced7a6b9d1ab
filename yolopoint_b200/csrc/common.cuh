// Shared helpers for libyolopoint_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/yolopoint_b200.h"

namespace yp {

void set_error(const char* fmt, ...);

#define YP_CUDA_OK(expr)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      yp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
      return YP_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define YP_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      yp::set_error(__VA_ARGS__);    \
      return (code);                 \
    }                                \
  } while (0)

#define YP_LAUNCH_OK()                                                                        \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      yp::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return YP_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int fmt_planes(int fmt) { return fmt == YP_FMT_F32X2 ? 2 : 1; }
static inline int fmt_esize(int fmt) { return fmt == YP_FMT_BF16 ? 2 : 4; }

int sm_count();

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float tf32_round(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// value of an activation element at element offset `off` of a view (plane 0 base pointer `base`)
__device__ __forceinline__ float load_act(const void* base, int fmt, int64_t plane_stride, int64_t off) {
  if (fmt == YP_FMT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[off]);
  const float* f = reinterpret_cast<const float*>(base);
  float v = f[off];
  if (fmt == YP_FMT_F32X2) v += f[off + plane_stride];
  return v;
}

__device__ __forceinline__ void store_act(void* base, int fmt, int64_t plane_stride, int64_t off, float v) {
  if (fmt == YP_FMT_BF16) {
    reinterpret_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(v);
  } else if (fmt == YP_FMT_F32X2) {
    float hi = tf32_round(v);
    float* f = reinterpret_cast<float*>(base);
    f[off] = hi;
    f[off + plane_stride] = tf32_round(v - hi);
  } else {
    reinterpret_cast<float*>(base)[off] = v;
  }
}

__device__ __forceinline__ float silu_accurate(float v) { return v / (1.0f + expf(-v)); }

}  // namespace yp
