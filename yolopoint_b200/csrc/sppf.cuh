// SPPF: three chained 5x5 stride-1 max pools (padding = -inf, i.e. clipped windows), all in shared memory
// (src/models/common.py:213-229 of the reference: y1 = m(x), y2 = m(y1), y3 = m(y2), cat(x, y1, y2, y3)).
// One work item owns G channels of one image: stage the (value, source pixel) pairs of the whole map, run the three
// pooling passes ping-pong between two shared buffers and copy the operand planes of the winning source pixel into
// concat slices 1..3, so the stored (hi, lo) pairs are bit-identical to the source element's.
// Shared by sppf_pool_kernel (layout.cu: one CTA per item) and conv_chain_kernel (conv_tc.cu: an operation of a layer chain).
#pragma once
#include "common.cuh"

namespace yp {

constexpr int SPPF_G = 2;   // channels per item: small groups -> C/2 items per image (the map is tiny, parallelism comes from channels)

static inline size_t sppf_smem_bytes(int HW) { return static_cast<size_t>(HW) * SPPF_G * (2 * sizeof(float) + 2 * sizeof(unsigned short)); }

// item = (channel group cg, image b); all threads of the CTA take part; sp_smem: sppf_smem_bytes(H*W) bytes, 4-byte aligned
__device__ __forceinline__ void sppf_pool_item(const YpView& cat4, int C, int cg, int b, unsigned char* sp_smem) {
  const int H = cat4.H, W = cat4.W, HW = H * W;
  float* val[2] = {reinterpret_cast<float*>(sp_smem), reinterpret_cast<float*>(sp_smem) + HW * SPPF_G};
  unsigned short* src[2] = {reinterpret_cast<unsigned short*>(val[1] + HW * SPPF_G), reinterpret_cast<unsigned short*>(val[1] + HW * SPPF_G) + HW * SPPF_G};
  const int c0 = cg * SPPF_G;
  const int64_t img = static_cast<int64_t>(b) * HW * cat4.pix_stride;
  const int n = HW * SPPF_G;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int p = i / SPPF_G, g = i - p * SPPF_G;
    val[0][i] = load_act(cat4.base, cat4.format, cat4.plane_stride, img + static_cast<int64_t>(p) * cat4.pix_stride + c0 + g);
    src[0][i] = static_cast<unsigned short>(p);
  }
  __syncthreads();
  for (int pass = 0; pass < 3; ++pass) {
    const float* vi = val[pass & 1];
    const unsigned short* si = src[pass & 1];
    float* vo = val[(pass + 1) & 1];
    unsigned short* so = src[(pass + 1) & 1];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int p = i / SPPF_G, g = i - p * SPPF_G;
      const int y = p / W, x = p - y * W;
      float best = -INFINITY;
      unsigned short arg = 0;
      for (int yy = max(0, y - 2); yy <= min(H - 1, y + 2); ++yy)
        for (int xx = max(0, x - 2); xx <= min(W - 1, x + 2); ++xx) {
          const int q = (yy * W + xx) * SPPF_G + g;
          const float v = vi[q];
          if (v > best) { best = v; arg = si[q]; }
        }
      vo[i] = best;
      so[i] = arg;
      // slice (pass+1) of the concat buffer <- planes of the winning source element
      const int64_t dst = img + static_cast<int64_t>(p) * cat4.pix_stride + c0 + g + static_cast<int64_t>(pass + 1) * C;
      const int64_t from = img + static_cast<int64_t>(arg) * cat4.pix_stride + c0 + g;
      if (cat4.format == YP_FMT_BF16) {
        __nv_bfloat16* f = static_cast<__nv_bfloat16*>(cat4.base);
        f[dst] = f[from];
      } else {
        float* f = static_cast<float*>(cat4.base);
        f[dst] = f[from];
        if (cat4.format == YP_FMT_F32X2) f[dst + cat4.plane_stride] = f[from + cat4.plane_stride];
      }
    }
    __syncthreads();
  }
}

}  // namespace yp
