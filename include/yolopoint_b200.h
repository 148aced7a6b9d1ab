/*
 * yolopoint_b200.h  --  C ABI of libyolopoint_b200.so (B200 / sm_100a only).
 *
 * The reference (UniBwTAS/YOLOPoint) is pure Python/PyTorch and has no FFI seam:
 * its "operator interface" for the hot path is a handful of Python call signatures.  Each entry point
 * below is what a ctypes binding of one of those Python functions needs; the reference function it
 * replaces is cited as file:line relative to the reference repository root.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library never allocates device
 *     memory, never frees and never keeps a pointer past the call;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises, so the
 *     calls compose with CUDA graphs; variable-length results are written into caller-sized buffers
 *     together with a device-side count;
 *   - return value: YP_OK (0) or a negative YpStatus; never throws across the ABI.
 *     yp_last_error() returns a thread-local message for the last failing call.
 *   - re-entrant; no global mutable state except a once-initialised driver entry point and function
 *     attributes.
 */
#ifndef YOLOPOINT_B200_H_
#define YOLOPOINT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YP_ABI_VERSION 9

typedef enum {
  YP_OK = 0,
  YP_ERR_SHAPE = -1,     /* unsupported / inconsistent shape */
  YP_ERR_ALIGN = -2,     /* pointer or stride not 16-byte aligned */
  YP_ERR_ARCH = -3,      /* device is not sm_100 */
  YP_ERR_CUDA = -4,      /* a CUDA call failed; see yp_last_error() */
  YP_ERR_CAPACITY = -5,  /* a caller-provided buffer is too small */
  YP_ERR_ARG = -6        /* invalid argument value (e.g. negative nn_thresh) */
} YpStatus;

int yp_abi_version(void);
const char* yp_last_error(void);
/* 0 if the current device can run this library (compute capability 10.x). */
int yp_check_device(void);
/* Stream-ordered copy between device and/or pinned host buffers (cudaMemcpyDefault); used by the host boundary of the whole-frame
 * pipeline (frame upload, count-bounded result read-back).  The reference moves the same data with tensor.to(device) /
 * .cpu().numpy() (src/demo.py:132, 138, 214). */
int yp_memcpy_async(void* dst, const void* src, size_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Activation views.  Activations live in HBM as NHWC tensors [planes][B][H][W][C_total]; a view is a
 * channel slice of one of them (that is how torch.cat / nn.Upsample of the reference disappear:
 * producers write into channel slices of their consumer's buffer).
 *   format YP_FMT_F32X2 : two fp32 planes (hi, lo), both pre-rounded to TF32, value = hi + lo
 *                         (operand format of the 3xTF32 tcgen05 path, fp32-grade accuracy)
 *          YP_FMT_BF16  : one bf16 plane (fast path)
 *          YP_FMT_F32   : one plain fp32 plane (network outputs: semi / desc / Detect logits)
 * ---------------------------------------------------------------------------------------------- */
typedef enum { YP_FMT_F32X2 = 0, YP_FMT_BF16 = 1, YP_FMT_F32 = 2 } YpFormat;

typedef struct {
  void* base;           /* element (plane 0, b 0, h 0, w 0, first channel of the slice) */
  int32_t B, H, W, C;   /* logical extent of the view; C = channels in the slice */
  int64_t pix_stride;   /* elements between consecutive pixels (= C_total of the buffer) */
  int64_t plane_stride; /* elements between planes (ignored for one-plane formats) */
  int32_t format;       /* YpFormat */
  int32_t upsample;     /* outputs only: 1 = store as is, 2 = store each pixel to the 2x2 block of a
                           (2H x 2W) destination (nn.Upsample(2,'nearest') fused into the producer);
                           H, W are then the SOURCE (pre-upsample) extents;
                           YP_UP_PARITY + 2*ph + pw = store pixel (h, w) to (2h+ph, 2w+pw) of a (2H x 2W) destination only
                           (one parity class of a stride-2 data gradient, see yp_conv2d_nhwc_fwd) */
} YpView;
#define YP_UP_PARITY 16

typedef enum { YP_ACT_NONE = 0, YP_ACT_SILU = 1 } YpAct;
typedef enum { YP_ALGO_TCGEN05 = 0, YP_ALGO_SIMT = 1 } YpConvAlgo;
#define YP_TILE_WIDE (-1)
#define YP_EPI_L2NORM 1u /* divide each output pixel by its L2 norm over all `cout` channels */
#define YP_EPI_NO_PATCH 2u /* planner hint: do not use the shared-memory patch formulation of 3x3 stride-1 convs */
#define YP_EPI_ROWMIN 4u   /* descriptor matching on the tensor cores: nothing is stored; for every input pixel i (a descriptor of set 1, the
                              1x1 "conv" weights being the descriptors of set 2) the epilogue reduces key(i, j) = float_bits(sqrt(2 - 2*clip(<d1_i, d2_j>,
                              -1, 1))) << 32 | (j + col_off) over the output channels j with an integer MIN into row_key[i] (the keys of
                              yp_match_partial; PointTracker.nn_match_two_way, demo.py:300-341).  Needs ksize 1, F32X2 operands, n_out = 0. */

/*
 * yp_conv2d_nhwc_fwd -- one Conv block of the reference in eval/fused form:
 *   act(conv(x) + bias) [+ residual] [-> L2 normalise], written to 1..2 destinations.
 * Replaces models/common.py:22-34 (Conv.forward_fuse, BN folded per utils/torch_utils_yolo.py:194-214),
 * the bare nn.Conv2d heads at models/YOLOPoint.py:186,195, Detect.m[i] at models/yolo.py:52, the residual
 * add of models/common.py:88, torch.cat/nn.Upsample at models/YOLOPoint.py:214-242 (via channel-slice and
 * 2x-replicated stores) and the descriptor normalisation at models/YOLOPoint.py:219-220.
 *
 * weight: packed [planes][cout][taps * Cin] in the input's operand format (F32X2 -> 2 fp32 planes hi/lo,
 *         BF16 -> 1 plane), K index = tap * Cin + cin, tap = kh * KW + kw.   bias: fp32 [cout] or NULL.
 * Supported geometry: 1x1 s1 p0, 3x3 s1 p1, 3x3 s2 p1 (the 6x6 s2 p2 stem is presented as 3x3 s1 p1 on
 * the 2x2 space-to-depth input produced by yp_frame_to_s2d / yp_nchw_to_s2d).
 * cout must be a multiple of 16; in.C a multiple of 16 (fp32) / 16 (bf16).
 */
typedef struct {
  YpView in;
  const void* weight;
  const float* bias;
  int32_t ksize, stride;  /* (1,1) (3,1) (3,2); (0,1) = custom tap list below */
  int32_t cout;
  int32_t act;            /* YpAct */
  uint32_t epilogue;      /* YP_EPI_* flags */
  YpView residual;        /* base == NULL -> none; same geometry/format family as out[0] */
  int32_t n_out;          /* 1 or 2 */
  YpView out[2];
  int32_t algo;           /* YpConvAlgo */
  int32_t tile_n;         /* 0 = library heuristic; YP_TILE_WIDE (-1) = throughput plan (widest tile, accumulators drained into registers so that the
                             fp32-grade accuracy holds at any width; split_k is then chosen by the library); -g (g >= 2) = the same with the persistent grid
                             capped at #SMs / g CTAs (a caller that keeps several frames in flight shares the GPU between their layers);
                             else the N (output-channel) tile: a divisor of cout, multiple of 16, <= 128 (fp32) / 256 (bf16) */
  int32_t split_k;        /* 0 = let the library slice K over several CTAs when the layer cannot fill the GPU, 1 = never, n = n slices */
  void* workspace;        /* split-K scratch (zero-initialised once by the caller, reusable by later launches on the same
                             stream; launches that may run concurrently need distinct workspaces); NULL -> never split */
  uint64_t workspace_bytes;
  /* ksize == 0: out[p] = sum_t W[:, t, :] . in[p + (tap_dh[t], tap_dw[t])], zero outside the input; 1 <= n_taps <= 9, offsets in
     [-8, 7].  Used for the data gradient of stride-2 convs (one launch per output parity class, see yolopoint_b200/train.py:
     autograd of models/common.py:22-34 in the reference's training step, train.py:208-220). */
  int32_t n_taps;
  int8_t tap_dh[9], tap_dw[9];
  /* YP_EPI_ROWMIN: */
  unsigned long long* row_key; /* [in.W] keys, pre-filled with ~0 by the caller */
  const int32_t* n_rows;       /* device count of valid input pixels (descriptors of set 1), NULL = in.W */
  const int32_t* n_cols;       /* device count of valid output channels (descriptors of set 2), NULL = cout */
  int32_t col_off;             /* added to j in the key (column shard offset of the multi-GPU match) */
  unsigned long long* col_key; /* optional [cout] keys (pre-filled with ~0): the epilogue also reduces key'(i, j) = dist_bits << 32 | i over the
                                  input pixels i with an integer MIN into col_key[j] -- the other direction of the two-way match from the same
                                  similarity tile (one pass instead of two) */
} YpConvDesc;

int yp_conv2d_nhwc_fwd(const YpConvDesc* desc, void* stream);
/* Bytes of split-K workspace yp_conv2d_nhwc_fwd would like for this descriptor (0 = it will not split). */
size_t yp_conv2d_workspace_bytes(const YpConvDesc* desc);
/* Dry run of yp_conv2d_nhwc_fwd's host side for this descriptor: format / shape / tile / shared-memory / TMEM planning without
 * touching the device or the pointers (works on a machine without a GPU).  Returns the status the launch would fail with before any
 * tensor map is encoded, YP_OK otherwise; yp_last_error() describes the failure.  Lets a caller validate a whole launch plan (every
 * layer of a model at a given input shape) ahead of time. */
int yp_conv2d_plan_check(const YpConvDesc* desc);

/*
 * Layer chains -- a segment of the network (the Conv / Bottleneck / C3 / SPPF sequence of YOLOPoint.forward,
 * models/YOLOPoint.py:198-246, models/common.py:22-34, 79-89, 124-135, 213-229) executed by ONE persistent kernel instead of one
 * launch per layer: one CTA per SM walks the operations in the given order; every operation is cut into the same work items a
 * stand-alone launch would run as CTAs (item i of an operation goes to CTA (offset + i) mod #CTAs, the offsets rotating so that
 * consecutive small layers land on different SMs), and an item starts as soon as the operations it depends on have published
 * their completion counters in global memory (release / acquire at GPU scope, proxy fences around the TMA transfers).  TMEM
 * allocation, tensor-map prefetch, launch latency and the grid-completion gap are paid once per segment, and independent layers
 * that are adjacent in the list (different branches of the network) run concurrently on different SMs.
 *
 * ops[i].deps lists the operations (indices < i of the same chain) whose output ops[i] reads or overwrites (read-after-write,
 * write-after-read and write-after-write on activation buffers); the caller derives them from its buffer plan
 * (yolopoint_b200/engine.py: chain_dependencies).  The order of `ops` must be a valid sequential order.  Split-K layers get their
 * own workspace inside the chain object (conv.workspace is ignored).  Chains longer than the kernel-parameter space allows are cut
 * into several kernels launched back to back.  Supported: YP_ALGO_TCGEN05 convolutions with F32X2 operands (the parity mode) and
 * the SPPF pooling; anything else returns YP_ERR_SHAPE and the caller keeps launching layer by layer.
 * A chain object may be launched on one stream at a time (its completion counters and workspaces are per object).
 */
typedef struct {
  int32_t type;        /* 0 = convolution (conv), 1 = SPPF pooling in place on the 4-slice concat buffer `pool` (see yp_sppf_pool) */
  YpConvDesc conv;
  YpView pool;
  int32_t n_deps;
  int32_t deps[8];
} YpChainOp;
int yp_conv_chain_create(const YpChainOp* ops, int32_t n_ops, void** chain);
int yp_conv_chain_launch(void* chain, void* stream);
int yp_conv_chain_destroy(void* chain);
/* Number of kernels yp_conv_chain_launch enqueues (1 unless the chain was cut), and the work items / shared memory of the chain. */
int yp_conv_chain_info(void* chain, int32_t* n_kernels, int32_t* n_items, int32_t* smem_bytes);
/* Debug aid (chains created while the environment variable YP_CHAIN_DEBUG is set): device pointer of kernel `kernel`'s timeline,
 * [n_ops][n_ctas][16] int64 %globaltimer stamps of the last item each CTA ran per operation (0 item start, 6 dependencies met,
 * 1 prologue done, 8 first TMA load issued, 9 first operands landed, 10 first k-block issued, 11 all MMAs issued, 2 accumulators
 * complete, 3 stores complete, 4 item end, 7 completion published); NULL when not recording. */
int yp_debug_conv_chain_timeline(void* chain, int32_t kernel, void** device_buf_i64, int32_t* n_ops, int32_t* n_ctas);

/*
 * yp_conv2d_nhwc_wgrad -- weight gradient of a Conv2d (bias-free, pad = k/2) for the training step: what autograd computes
 * for nn.Conv2d inside models/common.py:22-34 when train.py:208-220 calls loss.backward().
 *   dw[co][tap * Cin + ci] += sum_{b,oh,ow} dy[b,oh,ow,co] * x[b, oh*s + kh - k/2, ow*s + kw - k/2, ci]     (tap = kh*3 + kw)
 * x, dy: bf16 NHWC views; dw: fp32 [Cout][taps*Cin] (the packed layout of yp_conv2d_nhwc_fwd weights), ACCUMULATED into --
 * the caller zero-fills it (or keeps accumulating over micro-batches).  Geometry: 1x1 s1, 3x3 s1, 3x3 s2; Cin, Cout multiples
 * of 8; output width <= 254.  Partial sums of the CTAs that share a dw tile are combined with fp32 reductions in L2
 * (summation order, hence the last bits, may differ between runs -- like the reference's cuDNN wgrad).
 * The data gradient needs no entry point of its own: it is yp_conv2d_nhwc_fwd on dy with the transposed, tap-flipped weights
 * (stride 2: four launches with a custom tap list, one per output parity class).
 */
typedef struct {
  YpView x;
  YpView dy;
  int32_t ksize, stride;
  float* dw;
} YpWgradDesc;
int yp_conv2d_nhwc_wgrad(const YpWgradDesc* desc, void* stream);

/*
 * yp_bn_act_fwd / yp_bn_act_bwd -- training-mode BatchNorm2d (batch statistics) + SiLU of a Conv block, i.e. the `act(bn(.))` of
 * models/common.py:22-34 (BN eps / momentum from common.py:18-20) and its backward, on bf16 NHWC activations [P = B*H*W][C], C % 8 == 0.
 *   fwd: mean/var over the P pixels per channel -> out = act(gamma * (y - mean) * rstd + beta)  (act: 1 = SiLU, 0 = identity);
 *        running_mean / running_var (may be NULL) are updated in place with `momentum` and the unbiased variance like
 *        nn.BatchNorm2d; save [4][C] fp32 receives (mean, rstd, scale, shift) for the backward; acc [2][C] fp32 is scratch.
 *   bwd: dy = gamma*rstd * (dz - mean(dz) - xhat * mean(dz*xhat)), dz = dout * act'(z); dgamma_dbeta [2][C] fp32 receives
 *        (dbeta = sum dz, dgamma = sum dz*xhat).
 * Four streaming passes over the activation (2 + 4 bytes per element forward, 4 + 6 backward), fp32 reductions.
 */
int yp_bn_act_fwd(const void* y, int64_t P, int32_t C, const float* gamma, const float* beta, float* running_mean, float* running_var,
                  float momentum, float eps, int32_t act, void* out, float* save, float* acc, void* stream);
int yp_bn_act_bwd(const void* dout, const void* y, int64_t P, int32_t C, const float* gamma, const float* save, int32_t act, void* dy,
                  float* dgamma_dbeta, void* stream);

/* Debug aid: when set to a device buffer of 512 + 3 * 20000 int64, CTA (0,0) of every following tcgen05 conv launch records
 * clock64 stamps of its pipeline events in the first 512 slots and every CTA records (SM id, start ns, end ns) behind them
 * (see tools/conv_timeline.py); NULL switches it off.  Not thread safe. */
int yp_debug_conv_timeline(void* device_buf_i64);

/* fp32 [n] -> the (hi, lo) TF32 operand planes of the 3xTF32 path (hi = tf32(x), lo = tf32(x - hi), round to nearest away): what
 * yolopoint_b200.engine.split_tf32 computes with PyTorch ops, as one kernel (operand preparation of the tensor-core match). */
int yp_split_tf32(const float* src, int64_t n, float* hi, float* lo, void* stream);

/* desc / ||desc||_2 over the channels of every pixel of a plain fp32 (YP_FMT_F32) NHWC view, in place -- the descriptor normalisation of
 * models/YOLOPoint.py:219-220 for descriptor widths that do not fit one accumulator tile (version "x", D = 320); narrower heads have it
 * fused into the last convolution's epilogue (YP_EPI_L2NORM). */
int yp_l2norm_nhwc(const YpView* view, void* stream);

/* SPPF pooling: from slice 0 (C channels) of the [B,H,W,4C] concat buffer compute the 5x5, 9x9 and 13x13
 * stride-1 max pools (== three chained MaxPool2d(5,1,2), models/common.py:220-229) into slices 1..3. */
int yp_sppf_pool(const YpView* cat4, void* stream);

/* MaxPool2d(kernel 2, stride 2) from view `in` [B,H,W,C] to view `out` [B,H/2,W/2,C] of the same format (C % 8 == 0): the
 * `descA = MaxPool(xa)` branch of the YOLOPointv52 descriptor head (models/YOLOPoint.py:287, 311).  `out` may be a channel slice of
 * a concat buffer.  The operand planes of the winning element are copied unchanged. */
int yp_maxpool2x2(const YpView* in, const YpView* out, void* stream);

/* Input conversion.  Both produce the 2x2 space-to-depth NHWC operand of the stem conv:
 * out[b, h2, w2, (ph*2+pw)*3 + c] = x[b, c, 2*h2+ph, 2*w2+pw], channels 12..15 zero.
 *   yp_nchw_to_s2d : x fp32 NCHW [B,3,H,W]                       (Model.forward input, YOLOPoint.py:70)
 *   yp_frame_to_s2d: frame uint8 HWC [B,H,W,3], value/255         (demo.py:130-132 preprocessing)   */
int yp_nchw_to_s2d(const float* x, int32_t B, int32_t H, int32_t W, const YpView* out, void* stream);
int yp_frame_to_s2d(const uint8_t* frame, int32_t B, int32_t H, int32_t W, const YpView* out, void* stream);

/* NHWC view (any format) -> dense fp32 NCHW [B,C,H,W] (the layout Model.forward returns). */
int yp_nhwc_to_nchw(const YpView* in, int32_t C, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Detect decode (models/yolo.py:49-81, eval branch).
 * logits: fp32 NHWC [B,ny,nx,ldc] with channel a*no + o   ->
 *   raw  [B,na,ny,nx,no]  (the permuted logits the reference returns as x[i]; may be NULL)
 *   pred [B,A_total,no] rows [row_off, row_off + na*ny*nx): xy=(2s-0.5+grid)*stride, wh=(2s)^2*anchor, s
 * anchors_px: na*2 floats on the HOST = Detect.anchors[i] * stride[i] (pixels).
 * ---------------------------------------------------------------------------------------------- */
int yp_detect_decode(const float* logits, int32_t B, int32_t ny, int32_t nx, int32_t ldc, int32_t na, int32_t no,
                     float stride_px, const float* anchors_px_host, float* raw, float* pred, int64_t A_total,
                     int64_t row_off, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Box NMS = utils/general_yolo.py:124-235 (non_max_suppression) incl. torchvision.ops.nms semantics.
 * pred [B,A,no] fp32.  Per image: obj > conf_thres; conf = cls*obj; xywh->xyxy; multi_label -> one
 * candidate per (row, class) with conf > thr in row-major order, else best class; optional class filter
 * (class_mask: 32-bit words, bit c set = keep class c; NULL = all); stable sort by conf desc, cap max_nms;
 * greedy IoU (> iou_thres suppresses; boxes offset by cls*max_wh unless agnostic); cap max_det.
 * out_boxes [B,max_det,6] (x1,y1,x2,y2,conf,cls), out_count int32 [B].
 * workspace: yp_box_nms_workspace_bytes(B, A, no, cap) bytes (~(2A + 7 cap) * 4 B per image; no n x n matrix);
 * cap = per-image candidate capacity, a multiple of 64.  With cap >= max_nms (e.g. 30016 for the reference's 30000) ANY
 * number of candidates is handled the way the reference does: the max_nms most confident are kept (exact radix select,
 * ties in candidate order) and the call never reports overflow.  Only with cap < max_nms can a frame overflow; then
 * out_count[b] = -1 - n_candidates so that the caller can grow cap.  The first 4 int32 per image of the workspace hold
 * statistics of the last call: rows passing objectness, candidates found, candidates that entered the NMS, path
 * (0 = shared memory, 1 = workspace arrays, 2 = top-max_nms select, 3 = overflow reported).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float conf_thres, iou_thres;
  int32_t multi_label, agnostic, max_det, max_nms;
  float max_wh;
  const uint32_t* class_mask; /* device, (nc+31)/32 words, or NULL */
} YpNmsParams;

size_t yp_box_nms_workspace_bytes(int32_t B, int64_t A, int32_t no, int32_t cap);
/* Fused front end for the whole-frame pipeline: Detect decode (models/yolo.py:60-68) + the NMS above, straight from the
 * three levels' raw logits (fp32 NHWC [B,ny,nx,ldc], channel a*no+o), so that `pred` is never written: only rows whose
 * sigmoid(objectness) passes conf_thres are decoded.  Arrays of 3 (levels) live on the HOST; anchors_px_host = 18 floats
 * (level, anchor, w/h) in pixels.  Same outputs / workspace as yp_box_nms with A = sum_l na*ny*nx. */
int yp_detect_nms(const float* const* logits3_host, const int32_t* ny3_host, const int32_t* nx3_host, const int32_t* ldc3_host,
                  const float* stride3_host, const float* anchors_px_host, int32_t B, int32_t na, int32_t no, const YpNmsParams* p,
                  int32_t cap, float* out_boxes, int32_t* out_count, void* workspace, size_t workspace_bytes, int32_t prescanned_levels,
                  void* stream);
/* Optional: list the candidates of ONE level ahead of yp_detect_nms, e.g. on a side stream as soon as that level's Detect
 * convolution is done (levels 0 / 1 hold 95 % of the rows and finish long before level 2).  yp_detect_nms then gets the bit mask of
 * the levels listed this way (prescanned_levels, 0 = none) and only scans the others.  Same arguments / workspace as yp_detect_nms;
 * the workspace must be zero-filled once before the first use (the calls leave its counters zero again). */
int yp_detect_prescan(const float* const* logits3_host, const int32_t* ny3_host, const int32_t* nx3_host, const int32_t* ldc3_host,
                      const float* stride3_host, const float* anchors_px_host, int32_t B, int32_t na, int32_t no, const YpNmsParams* p,
                      int32_t cap, int32_t level, void* workspace, size_t workspace_bytes, void* stream);
int yp_box_nms(const float* pred, int32_t B, int64_t A, int32_t no, const YpNmsParams* p, int32_t cap,
               float* out_boxes, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Keypoint heatmap = utils/utils.py:232-262 (flattenDetection) / demo.py:140-150.
 * semi: logits, element (b,c,hc,wc) at b*sB + c*sC + hc*sH + wc*sW (so NCHW and NHWC both work).
 * heat [B, 8*Hc, 8*Wc] fp32: heat[8hc+i, 8wc+j] = softmax_c(semi)[8i+j]; channel 64 (dustbin) dropped.
 * variant 0: torch.softmax (max-subtracted).  variant 1: demo.py's exp(x)/(sum+1e-5).
 * The normalisation is one correctly rounded reciprocal of the sum per cell and a multiply per pixel (<= 1 ulp from the quotient; the
 * two layouts give bit-identical results).  Rows whose channels are contiguous and 16-byte aligned (sC == 1, sW % 4 == 0) take
 * vector loads.
 * ---------------------------------------------------------------------------------------------- */
int yp_heatmap(const float* semi, int32_t B, int32_t Hc, int32_t Wc, int64_t sB, int64_t sC, int64_t sH, int64_t sW,
               int32_t variant, float* heat, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Keypoints = utils/utils.py:465-485 (getPtsFromHeatmap) + 118-182 (nms_fast) == demo.py:151-166,
 * optionally followed by the keypoint-in-box filter of demo.py:178-198.
 * heat [B,H,W] fp32.  Candidates: heat >= conf_thresh.  Exact greedy Chebyshev-radius NMS (parallel
 * fixed point, identical survivors to the sequential reference), survivors sorted by confidence
 * descending (ties: reverse raster order), border filter x<border | x>=W-border | y<border | y>=H-border.
 * boxes (may be NULL): [B,box_ld,6] with box_count int32 [B]; points inside any box
 * (np.rint corners, half-open, python negative-index wrap) are dropped.
 * out_pts [B,max_pts,3] fp32 (x, y, conf), out_count int32 [B] (-1-n on overflow of max_pts).
 * ---------------------------------------------------------------------------------------------- */
size_t yp_keypoints_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t max_pts);
/* Whole-frame pipeline pieces (src/demo.py:151-198 split so that only the in-box filter waits for the box NMS):
 *   yp_keypoints_collect with boxes == NULL builds the confidence-ordered list of NMS survivors inside the border while the
 *   detection branch is still running; yp_keypoints_filter then removes the points inside the boxes as an order-preserving
 *   compaction: out_pts [B,max_pts,3], out_sel [B,max_pts] (index of each kept point in pts_all), out_count [B].  If row_key is
 *   given it also initialises the keys of the two-way match that follows (rows: n_prev[b] points of the previous frame, columns:
 *   the kept points).  yp_keypoints_threshold_count copies, per image, the number of pixels with heat >= conf_thresh seen by
 *   the last yp_keypoints_nms on this workspace (demo.py:152-153 returns early only when that number is 0). */
int yp_keypoints_filter(const float* pts_all, const int32_t* n_all, int32_t B, int32_t max_pts, int32_t H, int32_t W,
                        const float* boxes, const int32_t* box_count, int32_t box_ld, float* out_pts, int32_t* out_sel,
                        int32_t* out_count, const int32_t* n_prev, unsigned long long* row_key, unsigned long long* col_key,
                        void* stream);
int yp_keypoints_threshold_count(const void* workspace, size_t workspace_bytes, int32_t B, int32_t H, int32_t W, int32_t max_pts,
                                 int32_t* out_count, void* stream);
/* The two halves of yp_keypoints, so that a pipeline can run the NMS (which only needs the heatmap) on another stream
 * while the boxes are still being computed, and collect (border + in-box filters, sort, emit) afterwards. */
int yp_keypoints_nms(const float* heat, int32_t B, int32_t H, int32_t W, float conf_thresh, int32_t nms_dist, int32_t max_pts,
                     void* workspace, size_t workspace_bytes, void* stream);
int yp_keypoints_collect(const float* heat, int32_t B, int32_t H, int32_t W, int32_t border, const float* boxes,
                         const int32_t* box_count, int32_t box_ld, float* out_pts, int32_t* out_count, int32_t max_pts,
                         void* workspace, size_t workspace_bytes, void* stream);
int yp_keypoints(const float* heat, int32_t B, int32_t H, int32_t W, float conf_thresh, int32_t nms_dist,
                 int32_t border, const float* boxes, const int32_t* box_count, int32_t box_ld,
                 float* out_pts, int32_t* out_count, int32_t max_pts, void* workspace, size_t workspace_bytes,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * Descriptor sampling = demo.py:200-215 == evaluations/descriptor_evaluation.py:148-181:
 * bilinear grid_sample(align_corners=True, zeros padding) at pixel coords (x,y) of an image of size
 * (img_h,img_w), then L2 normalisation.  desc: element (b,d,hc,wc) at b*sB + d*sD + hc*sH + wc*sW.
 * pts [B,pts_ld,3] fp32 (x,y,conf), count int32 [B] (device) -> out [B,pts_ld,D] fp32 (row-major per point).
 * ---------------------------------------------------------------------------------------------- */
int yp_sample_desc(const float* desc, int32_t B, int32_t D, int32_t Hc, int32_t Wc, int64_t sB, int64_t sD,
                   int64_t sH, int64_t sW, int32_t img_h, int32_t img_w, const float* pts, const int32_t* count,
                   int32_t pts_ld, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Two-way nearest-neighbour match = demo.py:300-341 (PointTracker.nn_match_two_way).
 * d1 [n1_cap,D], d2 [n2_cap,D] fp32 row-major unit descriptors; n1/n2 device int32 counts (NULL -> caps).
 * dist = sqrt(2 - 2*clip(d1.d2,-1,1)); row argmin (first index on ties), col argmin, mutual + dist < thr.
 *   yp_match_partial: per-row / per-column packed keys (float_bits(dist) << 32 | index) for the column
 *     shard [col_off, col_off + n2) -- the cross-GPU reduction is an integer MIN over row keys
 *     (SURVEY.md section 8e); row_key uint64 [n1_cap], col_key uint64 [n2_cap].
 *   yp_match_finalize: row_key [n1], col_best_row int32 [n2_total] -> matches [n1_cap,3] fp32 rows
 *     (i, j, dist) ascending in i, match_count int32 [1].
 * ---------------------------------------------------------------------------------------------- */
int yp_match_partial(const float* d1, const int32_t* n1, int32_t n1_cap, const float* d2, const int32_t* n2,
                     int32_t n2_cap, int32_t D, int32_t col_off, unsigned long long* row_key,
                     unsigned long long* col_key, void* stream);
int yp_match_finalize(const unsigned long long* row_key, const int32_t* n1, int32_t n1_cap,
                      const unsigned long long* col_key, int32_t n2_total, float nn_thresh, float* matches,
                      int32_t* match_count, void* stream);
/* Whole-frame pipeline form of the two-way match (src/demo.py:300-341 between consecutive frames), all B images in two launches
 * whose grids do not depend on the buffer capacity: descriptor i of set 1 / 2 of image b is row sel1[b*cap+i] / sel2[b*cap+i] of
 * d1 / d2 (each [B,cap,D]; sel == NULL: row i), n1 / n2 [B] are device-side counts, row_key / col_key [B,cap] must hold ~0 in their
 * first n1 / n2 entries (yp_keypoints_filter does that).  matches [B,cap,3], match_count [B].  counts3 (optional, [3,B]): the
 * frame's result counts (kcount, bcount, match_count) gathered for a single read-back.  yp_gather_rows: dst[b,i,:] = src[b,sel[b,i],:]
 * for i < count[b] (the compact descriptor block the host reads back). */
int yp_match_frames(const float* d1, const int32_t* sel1, const int32_t* n1, const float* d2, const int32_t* sel2, const int32_t* n2,
                    int32_t B, int32_t cap, int32_t D, unsigned long long* row_key, unsigned long long* col_key, float nn_thresh,
                    float* matches, int32_t* match_count, const int32_t* kcount, const int32_t* bcount, int32_t* counts3, void* stream);
int yp_gather_rows(const float* src, const int32_t* sel, const int32_t* count, int32_t B, int32_t cap, int32_t D, float* dst, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Homography adaptation (export path).
 * yp_warp_image_batch = utils/utils.py:333-376 (warp_image_batch): img [B,C,H,W] fp32 -> out [B,C,H,W], sampled at the inverse
 * homography hinv [B,3,3] of every output pixel's normalised coordinate; xs [W] / ys [H] are torch.linspace(-1, 1, n) (passed in so
 * that the coordinates are bit-identical to the reference's); nearest = 0: bilinear, 1: nearest; align_corners = True, zeros padding.
 * yp_homography_adaptation = export_homography.py:97-128: heat, mask [B,H,W] (the B warped copies of ONE image and their valid
 * masks) -> agg [H,W] = sum_b warp(heat_b * mask_b) / sum_b warp(mask_b) in one pass (0/0 = NaN like the reference);
 * sum_heat / sum_mask [H,W] optionally receive numerator and denominator.
 * ---------------------------------------------------------------------------------------------- */
int yp_warp_image_batch(const float* img, const float* hinv, const float* xs, const float* ys, int32_t B, int32_t C, int32_t H, int32_t W,
                        int32_t nearest, float* out, void* stream);
int yp_homography_adaptation(const float* heat, const float* mask, const float* hinv, const float* xs, const float* ys, int32_t B,
                             int32_t H, int32_t W, float* sum_heat, float* sum_mask, float* agg, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Keypoint-detector loss of the training step, forward and backward fused:
 * labels2Dto3D + getMasks (utils/utils.py:184-209, 103-116) + ComputeDetectorLoss (utils/loss_functions.py:600-619).
 * semi: logits [B,65,Hc,Wc] fp32 with element strides (sB, sC, sH, sW); labels2d, mask2d: [B,8Hc,8Wc] fp32 contiguous.
 * dsemi (same strides as semi) receives d loss / d semi; out2[0] = loss, out2[1] = number of valid cells (sum of the cell mask).
 * workspace: yp_detector_loss_workspace_bytes(B, Hc, Wc).  Reductions run in a fixed order (bit-reproducible).
 * ---------------------------------------------------------------------------------------------- */
size_t yp_detector_loss_workspace_bytes(int32_t B, int32_t Hc, int32_t Wc);
int yp_detector_loss(const float* semi, int64_t sB, int64_t sC, int64_t sH, int64_t sW, const float* labels2d, const float* mask2d,
                     int32_t B, int32_t Hc, int32_t Wc, float* dsemi, float* out2, void* workspace, size_t workspace_bytes, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Object loss of the training step over all Detect levels, forward and gradient (csrc/object_loss.cu).
 * Replaces ComputeObjectLoss.__call__ + build_targets (utils/loss_functions.py:120-234: CIoU box loss via bbox_iou(..., CIoU=True) of
 * utils/metrics_yolo.py:202-240, BCE objectness against the detached, clamped CIoU, BCE classes) on the fixed-shape target plan
 * of ComputeObjectLoss.build_targets (:218-234; all 5 x anchors x targets assignment candidates with a validity mask).
 * Per level: pred / dpred [cells, no] fp32 rows (cells = B * na * ny * nx, row = (x, y, w, h, obj, classes...) logits);
 * valid [E] bytes, cell [E] flat (image, anchor, gj, gi) row index, tbox [E,4] target box relative to the cell, anchor [E,2] in cells,
 * cls [E]; balance = the level's objectness weight.  dpred is fully written (d loss / d pred).
 * out4 = (loss, box, obj, cls) with loss = w_box * box + w_obj * obj + w_cls * cls as the reference sums them before its batch-size factor.
 * ---------------------------------------------------------------------------------------------- */
#define YP_OBJ_LOSS_MAX_LEVELS 5
#define YP_OBJ_LOSS_MAX_ANCHORS 8
typedef struct YpObjLossLevel {
  const float* pred;
  float* dpred;
  const uint8_t* valid;
  const int64_t* cell;
  const float* tbox;
  const float* anchor;
  const int64_t* cls;
  int64_t cells;
  int32_t E;
  float balance;
  /* targets != NULL: the target assignment itself (build_targets, :218-234) is evaluated inside the kernels from the label list
   * targets [nt,6] = (image, class, x, y, w, h; box normalised to [0,1]) -- valid / cell / tbox / anchor / cls are then ignored and
   * E must be 5 * na * nt (candidate order: offset variant, anchor, target).  Grid nx x ny, nb images, anchors in cells. */
  const float* targets;
  int32_t nt, na, nx, ny, nb;
  float anchors[2 * YP_OBJ_LOSS_MAX_ANCHORS];
} YpObjLossLevel;

typedef struct YpObjLossParams {
  float cp, cn;          /* positive / negative class targets (label smoothing) */
  float cls_pw, obj_pw;  /* BCE positive-class weights */
  float gr;              /* objectness target = (1 - gr) + gr * max(CIoU, 0) */
  float w_box, w_obj, w_cls;
  float eps;             /* 1e-7 in the reference */
  float anchor_t;        /* target / anchor side-ratio bound of the assignment (4.0 in the reference's hyper-parameters) */
} YpObjLossParams;

size_t yp_object_loss_workspace_bytes(const YpObjLossLevel* levels, int32_t nl);
int yp_object_loss(const YpObjLossLevel* levels, int32_t nl, int32_t no, int32_t nc, const YpObjLossParams* hp, float* out4,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Training-step glue on bf16 NHWC activations (csrc/glue.cu), each with its backward pass.
 * yp_cat_nhwc_fwd / _bwd: torch.cat(parts, 1) (models/common.py:123-135, 151-165) with the resampling the network applies to a
 *   part on the way in -- nn.Upsample(scale_factor=2, "nearest") (models/YOLOPoint.py:214, 222-223) or nn.MaxPool2d(2, 2)
 *   (models/YOLOPoint.py:311).  out / dout: [B,H,W,sum C] contiguous; part i: src (and grad) [B,Hs,Ws,C_i] contiguous with
 *   (Hs, Ws) = (H, W) for YP_CAT_COPY, (H/2, W/2) for YP_CAT_UP2, (2H, 2W) for YP_CAT_POOL2.  Backward writes every non-null
 *   grad (YP_CAT_POOL2 also reads src: the gradient goes to the first maximum of each window in raster order, as max_pool2d's).
 * yp_sppf_train_fwd / _bwd: SPPF's cat(x, m(x), m(m(x)), m(m(m(x)))), m = MaxPool2d(5, 1, 2) (models/common.py:213-229).
 *   x / dx [B,H,W,C], out4 / dout4 [B,H,W,4C], arg [3][B][H*W][C] uint16 = the pixel of x every pooled value came from
 *   (written by the forward pass, read by the backward pass).  C % 8 == 0, H*W <= 65535.
 * ---------------------------------------------------------------------------------------------- */
#define YP_CAT_MAX_PARTS 4
#define YP_CAT_COPY 0
#define YP_CAT_UP2 1
#define YP_CAT_POOL2 2
typedef struct YpCatPart {
  const void* src;
  void* grad;
  int32_t C;
  int32_t mode;
} YpCatPart;
int yp_cat_nhwc_fwd(const YpCatPart* parts, int32_t n, void* out, int32_t B, int32_t H, int32_t W, void* stream);
int yp_cat_nhwc_bwd(const YpCatPart* parts, int32_t n, const void* dout, int32_t B, int32_t H, int32_t W, void* stream);
int yp_sppf_train_fwd(const void* x, void* out4, uint16_t* arg, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
int yp_sppf_train_bwd(const void* dout4, const uint16_t* arg, void* dx, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLOPOINT_B200_H_ */
