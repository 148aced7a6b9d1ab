// Plain fp32 SIMT direct convolution with the same contract as the tcgen05 kernel
// (yp_conv2d_nhwc_fwd, algo YP_ALGO_SIMT).  One thread per (pixel, output channel); no tiling.
// It exists to cross-check the tensor-core kernel on the device at full sizes and to bisect
// descriptor bugs; the model never selects it.
#include "common.cuh"

namespace yp {
namespace {

struct SimtArgs {
  YpView in, res, out[2];
  const void* w;
  const float* bias;
  int ksize, stride, cout, act, l2norm, n_out, Ho, Wo;
};

__device__ __forceinline__ float simt_value(const SimtArgs& a, int b, int oh, int ow, int co) {
  const int pad = a.ksize / 2;
  const int Cin = a.in.C;
  const int Ktot = a.ksize * a.ksize * Cin;
  const int wfmt = a.in.format;  // weights are packed in the input's operand format
  const int64_t wplane = static_cast<int64_t>(a.cout) * Ktot;
  float acc = 0.0f;
  for (int kh = 0; kh < a.ksize; ++kh) {
    const int ih = oh * a.stride + kh - pad;
    if (ih < 0 || ih >= a.in.H) continue;
    for (int kw = 0; kw < a.ksize; ++kw) {
      const int iw = ow * a.stride + kw - pad;
      if (iw < 0 || iw >= a.in.W) continue;
      const int64_t ioff = ((static_cast<int64_t>(b) * a.in.H + ih) * a.in.W + iw) * a.in.pix_stride;
      const int64_t woff = static_cast<int64_t>(co) * Ktot + (kh * a.ksize + kw) * Cin;
      for (int c = 0; c < Cin; ++c)
        acc = fmaf(load_act(a.in.base, a.in.format, a.in.plane_stride, ioff + c), load_act(a.w, wfmt, wplane, woff + c), acc);
    }
  }
  float v = acc + (a.bias ? a.bias[co] : 0.0f);
  if (a.act == YP_ACT_SILU) v = silu_accurate(v);
  if (a.res.base) {
    const int64_t roff = ((static_cast<int64_t>(b) * a.Ho + oh) * a.Wo + ow) * a.res.pix_stride + co;
    v += load_act(a.res.base, a.res.format, a.res.plane_stride, roff);
  }
  return v;
}

__device__ __forceinline__ void simt_store(const SimtArgs& a, int b, int oh, int ow, int co, float v) {
  for (int i = 0; i < a.n_out; ++i) {
    const YpView& o = a.out[i];
    if (o.upsample == 2) {
      const int H2 = 2 * a.Ho, W2 = 2 * a.Wo;
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const int64_t off = ((static_cast<int64_t>(b) * H2 + 2 * oh + ph) * W2 + 2 * ow + pw) * o.pix_stride + co;
          store_act(o.base, o.format, o.plane_stride, off, v);
        }
    } else {
      const int64_t off = ((static_cast<int64_t>(b) * a.Ho + oh) * a.Wo + ow) * o.pix_stride + co;
      store_act(o.base, o.format, o.plane_stride, off, v);
    }
  }
}

__global__ void conv_simt_kernel(const SimtArgs a, int64_t total) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int co = static_cast<int>(idx % a.cout);
  int64_t p = idx / a.cout;
  const int ow = static_cast<int>(p % a.Wo); p /= a.Wo;
  const int oh = static_cast<int>(p % a.Ho);
  const int b = static_cast<int>(p / a.Ho);
  simt_store(a, b, oh, ow, co, simt_value(a, b, oh, ow, co));
}

// L2-norm variant: one thread per pixel computes all channels twice (sum of squares, then scaled store).
__global__ void conv_simt_l2_kernel(const SimtArgs a, int64_t pixels) {
  const int64_t p0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (p0 >= pixels) return;
  int64_t p = p0;
  const int ow = static_cast<int>(p % a.Wo); p /= a.Wo;
  const int oh = static_cast<int>(p % a.Ho);
  const int b = static_cast<int>(p / a.Ho);
  float ss = 0.0f;
  for (int co = 0; co < a.cout; ++co) { const float v = simt_value(a, b, oh, ow, co); ss += v * v; }
  const float nrm = sqrtf(ss);
  for (int co = 0; co < a.cout; ++co) simt_store(a, b, oh, ow, co, simt_value(a, b, oh, ow, co) / nrm);
}

}  // namespace

int conv_simt_forward(const YpConvDesc& d, cudaStream_t st) {
  YP_REQUIRE((d.ksize == 1 && d.stride == 1) || (d.ksize == 3 && (d.stride == 1 || d.stride == 2)), YP_ERR_SHAPE,
             "conv(simt): k=%d s=%d unsupported", d.ksize, d.stride);
  YP_REQUIRE(d.n_out >= 1 && d.n_out <= 2, YP_ERR_SHAPE, "conv(simt): n_out=%d", d.n_out);
  SimtArgs a;
  a.in = d.in; a.res = d.residual; a.out[0] = d.out[0]; a.out[1] = d.out[d.n_out > 1 ? 1 : 0];
  a.w = d.weight; a.bias = d.bias; a.ksize = d.ksize; a.stride = d.stride; a.cout = d.cout; a.act = d.act;
  a.l2norm = (d.epilogue & YP_EPI_L2NORM) ? 1 : 0; a.n_out = d.n_out;
  a.Ho = d.in.H / d.stride; a.Wo = d.in.W / d.stride;
  const int64_t pixels = static_cast<int64_t>(d.in.B) * a.Ho * a.Wo;
  if (a.l2norm) {
    conv_simt_l2_kernel<<<static_cast<unsigned>(ceil_div64(pixels, 128)), 128, 0, st>>>(a, pixels);
  } else {
    const int64_t total = pixels * d.cout;
    conv_simt_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, st>>>(a, total);
  }
  YP_LAUNCH_OK();
  return YP_OK;
}

}  // namespace yp
