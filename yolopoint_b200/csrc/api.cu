// C-ABI glue: error reporting, device checks and the convolution dispatcher.
#include <stdarg.h>

#include "common.cuh"

namespace yp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

int conv_tc_forward(const YpConvDesc& d, cudaStream_t st);
size_t conv_tc_workspace_bytes(const YpConvDesc& d);
int conv_tc_plan_check(const YpConvDesc& d);
int conv_chain_create(const YpChainOp* ops, int n_ops, void** out);
int conv_chain_launch(void* chain, cudaStream_t st);
int conv_chain_destroy(void* chain);
int conv_chain_info(void* chain, int* n_kernels, int* n_items, int* smem);
int conv_chain_debug(void* chain, int k, void** dbg, int* n_ops, int* n_ctas);
int conv_simt_forward(const YpConvDesc& d, cudaStream_t st);
void set_conv_timeline(long long* p);
int wgrad_tc(const YpWgradDesc& d, cudaStream_t st);

}  // namespace yp

extern "C" int yp_abi_version(void) { return YP_ABI_VERSION; }

extern "C" const char* yp_last_error(void) { return yp::g_err; }

extern "C" int yp_check_device(void) {
  int dev = 0;
  YP_CUDA_OK(cudaGetDevice(&dev));
  int major = 0;
  YP_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  YP_REQUIRE(major == 10, YP_ERR_ARCH, "libyolopoint_b200 is built for sm_100a only; device %d has compute capability %d.x", dev, major);
  return YP_OK;
}

extern "C" int yp_conv2d_nhwc_fwd(const YpConvDesc* d, void* stream) {
  YP_REQUIRE(d && d->in.base && d->weight, YP_ERR_ARG, "conv: null pointer in descriptor");
  if (d->epilogue & YP_EPI_ROWMIN)
    YP_REQUIRE(d->row_key && d->n_out == 0 && d->ksize == 1 && d->algo == YP_ALGO_TCGEN05 && d->in.format == YP_FMT_F32X2, YP_ERR_ARG,
               "conv: YP_EPI_ROWMIN needs row_key, n_out = 0, ksize 1, the tcgen05 algorithm and F32X2 operands");
  else
    YP_REQUIRE(d->n_out >= 1 && d->out[0].base, YP_ERR_ARG, "conv: null output in descriptor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->algo == YP_ALGO_SIMT) return yp::conv_simt_forward(*d, st);
  if (d->algo == YP_ALGO_TCGEN05) return yp::conv_tc_forward(*d, st);
  yp::set_error("conv: unknown algo %d", d->algo);
  return YP_ERR_ARG;
}

extern "C" int yp_debug_conv_timeline(void* device_buf_i64) {
  yp::set_conv_timeline(static_cast<long long*>(device_buf_i64));
  return YP_OK;
}

extern "C" size_t yp_conv2d_workspace_bytes(const YpConvDesc* d) {
  if (!d || d->algo != YP_ALGO_TCGEN05) return 0;
  return yp::conv_tc_workspace_bytes(*d);
}

extern "C" int yp_conv2d_plan_check(const YpConvDesc* d) {
  YP_REQUIRE(d, YP_ERR_ARG, "conv: null descriptor");
  YP_REQUIRE(d->algo == YP_ALGO_TCGEN05, YP_ERR_ARG, "conv: plan check is defined for the tcgen05 algorithm only (algo %d)", d->algo);
  return yp::conv_tc_plan_check(*d);
}

extern "C" int yp_conv_chain_create(const YpChainOp* ops, int32_t n_ops, void** chain) {
  YP_REQUIRE(ops && chain && n_ops >= 1, YP_ERR_ARG, "chain: null pointer or empty operation list");
  *chain = nullptr;
  return yp::conv_chain_create(ops, n_ops, chain);
}

extern "C" int yp_conv_chain_launch(void* chain, void* stream) {
  YP_REQUIRE(chain, YP_ERR_ARG, "chain: null handle");
  return yp::conv_chain_launch(chain, static_cast<cudaStream_t>(stream));
}

extern "C" int yp_conv_chain_destroy(void* chain) {
  if (!chain) return YP_OK;
  return yp::conv_chain_destroy(chain);
}

extern "C" int yp_conv_chain_info(void* chain, int32_t* n_kernels, int32_t* n_items, int32_t* smem_bytes) {
  YP_REQUIRE(chain, YP_ERR_ARG, "chain: null handle");
  return yp::conv_chain_info(chain, n_kernels, n_items, smem_bytes);
}

extern "C" int yp_debug_conv_chain_timeline(void* chain, int32_t kernel, void** device_buf_i64, int32_t* n_ops, int32_t* n_ctas) {
  YP_REQUIRE(chain && device_buf_i64 && n_ops && n_ctas, YP_ERR_ARG, "chain: null pointer");
  return yp::conv_chain_debug(chain, kernel, device_buf_i64, n_ops, n_ctas);
}

extern "C" int yp_conv2d_nhwc_wgrad(const YpWgradDesc* d, void* stream) {
  YP_REQUIRE(d && d->x.base && d->dy.base && d->dw, YP_ERR_ARG, "wgrad: null pointer in descriptor");
  return yp::wgrad_tc(*d, static_cast<cudaStream_t>(stream));
}

// Stream-ordered copy between any two of {device, pinned host} buffers (cudaMemcpyDefault): the host boundary of the whole-frame
// pipeline issues its H2D / D2H transfers through this so that a frame costs a handful of driver calls, not framework dispatches.
extern "C" int yp_memcpy_async(void* dst, const void* src, size_t bytes, void* stream) {
  YP_REQUIRE(dst && src, YP_ERR_ARG, "memcpy_async: null pointer");
  if (bytes == 0) return YP_OK;
  YP_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return YP_OK;
}
