/*
 * ifcb_b200.h -- C ABI of the B200-native hot path of WHOIGit/ifcb_classifier.
 *
 * The reference is pure Python and exposes no C interface; its only "operator
 * interfaces" on this path are Dataset.__getitem__ and nn.Module.forward.  Each
 * entry point below names the reference code it replaces (file:line relative to
 * the upstream repository, v0.3.1) -- INTEGRATION.md shows the ctypes binding a
 * maintainer would add on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the ABI.
 *   - Return value: 0 = ok, <0 = argument / shape error, >0 = cudaError_t.
 *     ifcb_last_error() returns a thread-local description of the last failure.
 *   - All device buffers are CALLER-OWNED (the Python host passes torch tensor
 *     storage); the library never allocates or frees device memory on the hot
 *     path.  Pointers prefixed d_ are device pointers, h_ are host pointers.
 *   - Every call is asynchronous on the cudaStream_t passed as `stream`
 *     (a void* so that this header needs no CUDA include).
 *   - Activations are NHWC 16-bit floats (bf16 or fp16, one format per plan:
 *     every layer descriptor carries `dtype` = IFCB_ACT_*); a tensor view is (base pointer, pixel stride
 *     `ld` in elements, channels) so that a layer can read or write a channel
 *     slice of a wider (concatenated) tensor in place.
 */
#ifndef IFCB_B200_H
#define IFCB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IFCB_B200_ABI_VERSION 1

int ifcb_abi_version(void);
const char* ifcb_last_error(void);
/* Number of SMs of the current device (148 on B200); <0 on error. */
int ifcb_sm_count(void);

/* Host-side .adc parser (no GPU work): the ROI table of a bin from its .adc file image -- what pyifcb's
 * bin.images / bin.schema expose upstream (reference neuston_data.py:446-454; pandas CSV parse there).
 *   buf/len      the file contents (headerless CSV, one row per trigger)
 *   col_*        0-based column of ROI_WIDTH, ROI_HEIGHT, START_BYTE (schema v2: 15, 16, 17; v1: 11, 12, 13)
 *   outputs      host arrays with room for max_rows entries; rows with zero area are dropped; targets are
 *                1-based row numbers.  Returns the number of ROIs kept, or -1 (ifcb_last_error). */
int64_t ifcb_parse_adc(const char* buf, int64_t len, int col_w, int col_h, int col_b, int64_t max_rows,
                       int32_t* targets, int64_t* offsets, int32_t* heights, int32_t* widths);

/* Host-side: JSON text "[[s00, s01, ...], [s10, ...], ...]" of a float32 [rows, cols] score matrix, byte-identical to
 * json.dumps(scores.tolist()) as the reference's .json result files hold it (neuston_callbacks.py:213-230; float32
 * widened to float64, shortest round-trip digits).  out must hold rows*(cols*28+4)+4 bytes; returns the text length
 * (no terminator) or -1. */
int64_t ifcb_format_scores_json(const float* scores, int64_t rows, int64_t cols, char* out, int64_t cap);

/* ------------------------------------------------------------------------- *
 * K1  fused ROI preprocess.
 * Replaces IfcbBinDataset.__getitem__ (neuston_data.py:456-464):
 *   ToPILImage('L') -> convert('RGB') -> Resize((R,R)) -> ToTensor() -> [Normalize]
 * for all ROIs of a bin in one launch, reading the raw .roi file image
 * (neuston_data.py:446-454 / pyifcb bin.images) straight from device memory.
 * Bit-exact restatement of Pillow's fixed-point antialiased bilinear resample.
 *
 *   d_packed      raw ROI bytes (the .roi file contents), `packed_bytes` long
 *   d_offsets[n]  START_BYTE of each ROI (int64)
 *   d_h, d_w[n]   ROI height / width (int32); ROI i is row-major (h, w) u8
 *   max_h, max_w  host-known upper bounds of d_h / d_w (the IFCB camera frame is
 *                 1034 x 1380); they size the launch's shared memory (any bound up
 *                 to several thousand pixels is accepted: very large images read
 *                 their horizontal pass from global memory)
 *   d_status      optional device word (NULL to skip).  A ROI whose table entry
 *                 points outside [0, packed_bytes) or whose size exceeds what the
 *                 launch was sized for is NOT processed: its output slot is zeroed
 *                 and IFCB_PRE_BAD_TABLE / IFCB_PRE_TOO_LARGE is OR-ed into the word
 *   R             output side (299 for inception_v3, 224 otherwise)
 *   h_mean,h_std  3 floats each (--img-norm, neuston_data.py:331-339) or NULL
 *   out_mode      IFCB_OUT_*; d_out holds n images of that layout
 *   pass_rule     IFCB_PASS_PILLOW12: Pillow >= 12 pass order (vertical pass
 *                 first when h > R and h > 100*w), IFCB_PASS_HV: Pillow 8.4.0
 *                 (the reference's pin; always horizontal first)
 * ------------------------------------------------------------------------- */
enum {
  IFCB_OUT_F32_NCHW = 0,  /* float32 [n,3,R,R]  == the reference's item tensor   */
  IFCB_OUT_BF16_NCHW = 1, /* bfloat16 [n,3,R,R]                                   */
  IFCB_OUT_U8_GRAY = 2    /* uint8 [n,R,R] resized gray plane (feeds the stem)    */
};
enum { IFCB_PASS_PILLOW12 = 0, IFCB_PASS_HV = 1 };
enum { IFCB_PRE_BAD_TABLE = 1, IFCB_PRE_TOO_LARGE = 2 };

int ifcb_preprocess(const uint8_t* d_packed, int64_t packed_bytes,
                    const int64_t* d_offsets, const int32_t* d_h, const int32_t* d_w,
                    int n, int max_h, int max_w, int R,
                    const float* h_mean, const float* h_std,
                    int out_mode, void* d_out, int pass_rule, int32_t* d_status, void* stream);

/* ------------------------------------------------------------------------- *
 * Network plan: an ordered list of layer launches over caller-owned buffers.
 * Replaces NeustonModel.forward + test_step (neuston_models.py:66-68,152-157),
 * i.e. the torchvision graph built by get_namebrand_model (neuston_models.py:22-45)
 * run in eval mode followed by softmax(dim=1), and the argmax/max of
 * save_run_results (neuston_callbacks.py:161-162).
 * ------------------------------------------------------------------------- */
typedef struct ifcb_plan ifcb_plan;

/* 16-bit storage / tensor-core operand format of activations and conv weights.
 * Accumulation is always fp32 (TMEM); BN scale/shift, the stem and the head are fp32. */
enum { IFCB_ACT_BF16 = 0, IFCB_ACT_FP16 = 1 };

int ifcb_plan_create(ifcb_plan** out);
int ifcb_plan_destroy(ifcb_plan* plan);
/* Launches every layer for `batch` images (batch <= the capacity the layers
 * were created with) on `stream`. */
int ifcb_plan_run(ifcb_plan* plan, int batch, void* stream);
/* Same, but the head writes its rows at `out_row` of the score / logit / top-1 buffers (which the caller
 * allocated with >= out_row + batch rows): the batches of one bin land side by side, so test_epoch_end's
 * concatenation (neuston_models.py:159-180) needs no copy and a bin leaves the device in ONE transfer. */
int ifcb_plan_run_at(ifcb_plan* plan, int batch, int out_row, void* stream);
/* Runs layers [first, last) only (layer-level tests, profiling). */
int ifcb_plan_run_range(ifcb_plan* plan, int first, int last, int batch, void* stream);
/* Call after rewriting a plan's weights / folded BN vectors in place (same buffers, e.g. the current weights of a TRAIN run
 * before each validation pass, neuston_models.py:94-103): refreshes the values the plan caches on the host.  Synchronous. */
int ifcb_plan_refresh(ifcb_plan* plan);
int ifcb_plan_num_layers(const ifcb_plan* plan);
/* Kernel launches one ifcb_plan_run performs. */
int ifcb_plan_num_launches(const ifcb_plan* plan);

/* One output segment of a (possibly horizontally fused) convolution: GEMM
 * columns [n_begin, n_end) go to d_out + pixel*ld + (n - n_begin). */
typedef struct {
  int32_t n_begin, n_end; /* multiples of 16 */
  void* d_out;            /* bf16, channel-slice base pointer */
  int32_t ld;             /* pixel stride of the destination, elements */
  int32_t relu;           /* 1: max(0, .) after the affine */
  int32_t pad_h, pad_w;   /* physical zero border of the destination tensor: it is
                             [batch, P+2*pad_h, Q+2*pad_w, ld] and d_out points at its
                             first (border) pixel; only interior pixels are written */
} ifcb_conv_segment;

#define IFCB_MAX_SEGMENTS 4
enum { IFCB_CONV_AUTO = 0, IFCB_CONV_IM2COL = 1, IFCB_CONV_WINDOW = 2, IFCB_CONV_IM2COL_PAIR = 3 };

/* K2  Conv2d(bias=False) + folded BatchNorm + ReLU as a tcgen05 implicit GEMM
 * (torchvision BasicConv2d, inception.py:398-407; ResNet conv-bn-relu).
 * y[n, co] = act( scale[co] * sum_{r,s,ci} x[n, r, s, ci] * w[co, r, s, ci] + shift[co] (+ residual) )
 *   d_in        NHWC view with pixel stride in_ld; logical extent [batch_cap, H, W, Cin],
 *               stored with a zero border of in_pad_h / in_pad_w pixels (physical
 *               [batch_cap, H+2*in_pad_h, W+2*in_pad_w, in_ld]); d_in = first border pixel
 *   algo        IFCB_CONV_IM2COL: one TMA im2col load per filter tap (any stride);
 *               IFCB_CONV_IM2COL_PAIR: the same on CTA pairs (cluster of 2, tcgen05
 *               cta_group::2): 256-pixel tiles, each SM loads half of every weight tile;
 *               IFCB_CONV_WINDOW: stride 1 and in_pad == pad -- the padded input patch
 *               of a tile is loaded ONCE and filter taps are shifted UMMA descriptors (the
 *               library runs it on CTA pairs when the layer has >= 64 output channels);
 *               IFCB_CONV_AUTO picks per shape (ifcb_conv_auto_config) among those that apply
 *   d_residual  view with the output's logical extent and its own zero border res_pad_*
 *   d_weight    16-bit [Cout_pad, kh*kw*Cin_pad] packed by the host: K index =
 *               (r*kw + s)*Cin_pad + ci, zero filled; Cin_pad = 32 for Cin <= 32 (64-byte
 *               operand rows), else round_up(Cin, 64); Cout_pad = n_tiles * tile_n
 *               (ifcb_conv_geometry returns all four)
 *   d_scale/d_shift  float32[Cout_pad] folded BN (gamma/sqrt(var+eps), beta - mean*that)
 *   d_residual  optional bf16 view added before the activation (ResNet), or NULL
 *   tile_n      GEMM N tile (multiple of 16, 16..256); 0 = library picks
 *   d_stats     see the field: per-channel sum / sum of squares of the stored outputs, accumulated by the epilogue
 */
typedef struct {
  const void* d_in;
  int32_t in_ld, Cin;
  int32_t batch_cap, H, W;
  int32_t in_pad_h, in_pad_w;
  int32_t kh, kw, stride_h, stride_w, pad_h, pad_w;
  int32_t Cout;
  const void* d_weight;
  const float* d_scale;
  const float* d_shift;
  int32_t n_seg;
  ifcb_conv_segment seg[IFCB_MAX_SEGMENTS];
  const void* d_residual;
  int32_t res_ld, res_pad_h, res_pad_w;
  int32_t tile_n;
  int32_t algo;  /* IFCB_CONV_* */
  int32_t dtype; /* IFCB_ACT_* of input, weights, residual and outputs */
  double* d_stats; /* optional (TRAIN, one segment): float64 [2][Cout] accumulators.  The epilogue ADDS, per output
                      channel, the sum and the sum of squares of the 16-bit values it stores (sum over every valid
                      output pixel of the launch) -- BatchNorm's batch statistics without a pass over the tensor
                      (torch.nn.BatchNorm2d in train mode, reached from neuston_models.py:80-86).  NULL: off */
} ifcb_conv_desc;

int ifcb_plan_add_conv(ifcb_plan* plan, const ifcb_conv_desc* desc);
/* The algorithm (IFCB_CONV_IM2COL / IFCB_CONV_WINDOW) and N tile the library prefers for a
 * conv over an H x W input -- what IFCB_CONV_AUTO / tile_n = 0 resolve to when the input
 * buffer carries the padding WINDOW needs.  The host asks BEFORE allocating the producer's
 * output so that only WINDOW inputs get a zero border, and passes tile_n on to
 * ifcb_conv_geometry for weight packing. */
int ifcb_conv_auto_config(int H, int W, int Cin, int Cout, int kh, int kw, int stride_h, int stride_w,
                          int pad_h, int pad_w, int32_t* algo, int32_t* tile_n);
/* N tile the library uses for `Cout` output channels under a given (non-AUTO) algorithm. */
int ifcb_conv_auto_tile_n(int Cout, int algo);
/* Packed-weight geometry for a conv: Cin_pad, K_pad = kh*kw*Cin_pad, tile_n and
 * Cout_pad the library will use (host packs weights / scale / shift to these). */
int ifcb_conv_geometry(int Cin, int Cout, int kh, int kw, int tile_n_hint,
                       int32_t* Cin_pad, int32_t* K_pad, int32_t* tile_n, int32_t* Cout_pad);

/* ------------------------------------------------------------------------- *
 * TRAIN step primitives.  Replace what loss.backward() + optimizer.step() run for
 * NeustonModel.training_step (neuston_models.py:70-86, configure_optimizers :63-64):
 * torch autograd's conv / batch-norm / pooling backward kernels and Adam.
 * ------------------------------------------------------------------------- */

/* Conv2d weight gradient on the tensor cores (pixel axis = GEMM K, MN-major operands):
 *   dW[co, r, s, ci] += sum_{n,p,q} dout[n,p,q,co] * in[n, p*stride_h + r - pad_h, q*stride_w + s - pad_w, ci]
 *   d_in       NHWC 16-bit view, logical extent [batch, H, W, Cin], pixel stride in_ld, stored with a
 *              zero border of in_pad_h / in_pad_w pixels (d_in = first border pixel), as ifcb_conv_desc
 *   d_dout     NHWC 16-bit view, logical extent [batch, P, Q, Cout], pixel stride dout_ld, zero border dout_pad_*
 *   d_dweight  float32 [Cout, kh*kw, Cin]; the kernel ACCUMULATES (split-K over CTAs with
 *              red.global.add.f32): zero it first for a plain gradient
 */
typedef struct {
  const void* d_in;
  int32_t in_ld, Cin;
  int32_t batch, H, W;
  int32_t kh, kw, stride_h, stride_w, pad_h, pad_w;
  const void* d_dout;
  int32_t dout_ld, Cout;
  float* d_dweight;
  int32_t dtype; /* IFCB_ACT_* of d_in and d_dout */
  int32_t in_pad_h, in_pad_w;
  int32_t dout_pad_h, dout_pad_w; /* zero border of the d_dout tensor (d_dout = its first border pixel) */
} ifcb_wgrad_desc;
int ifcb_conv_wgrad(const ifcb_wgrad_desc* desc, void* stream);
/* Deterministic TRAIN mode (the reference runs Trainer(deterministic=True), neuston_net.py:101): while a caller-owned device
 * workspace is set, ifcb_conv_wgrad STORES its split-K partial sums there and adds them into d_dweight in split order, and the
 * BatchNorm reductions (ifcb_bn_stats / ifcb_bn_backward) store per-block partials and sum them in block order -- no
 * floating-point atomics on the parameter path: a step is bitwise reproducible.  Process-wide; NULL / 0 switches it off.
 * ifcb_conv_wgrad_workspace_bytes: what one layer needs (the BatchNorm partials need < 2 MB). */
int ifcb_train_deterministic(void* d_workspace, int64_t bytes);
int64_t ifcb_conv_wgrad_workspace_bytes(const ifcb_wgrad_desc* desc);

/* A channel slice of an NHWC 16-bit activation (or gradient) tensor, as the Python host's
 * graph.View: logical extent [batch, H, W, C], stored with a zero border of pad_h / pad_w pixels
 * (physical [batch, H+2*pad_h, W+2*pad_w, ld]); d = first (border) pixel of the slice.
 * Kernels only ever write interior pixels. */
typedef struct {
  void* d;
  int32_t ld, C;
  int32_t H, W;
  int32_t pad_h, pad_w;
} ifcb_view;

/* cudaMemsetAsync(d, 0, bytes) on `stream` (gradient arena / accumulators at step start). */
int ifcb_memset_zero(void* d, int64_t bytes, void* stream);

/* BatchNorm2d, train mode (torch.nn.BatchNorm2d.forward with track_running_stats: batch mean and
 * BIASED variance normalise; running_mean/var updated with `momentum`, running_var with the
 * UNBIASED variance).  Split in two launches so that the conv output z is read twice and the
 * activation written once:
 *   ifcb_bn_stats   z -> d_mean[C], d_invstd[C] = 1/sqrt(var+eps)  (+ running stats when non-NULL)
 *                   d_acc: 2*C float64 accumulators at the start of a 64 KB scratch (8192 float64), zero on entry,
 *                   zero again on exit
 *   ifcb_bn_apply   out = act(gamma*(z-mean)*invstd + beta (+ residual)); relu: 0/1
 *                   (ResNet: bn3 + identity + ReLU, resnet.py Bottleneck/BasicBlock.forward) */
int ifcb_bn_stats(const ifcb_view* z, int batch, int dtype, float eps, float momentum, double* d_acc,
                  float* d_mean, float* d_invstd, float* d_running_mean, float* d_running_var, void* stream);
int ifcb_bn_apply(const ifcb_view* z, const ifcb_view* out, const ifcb_view* residual, int batch, int dtype,
                  const float* d_mean, const float* d_invstd, const float* d_gamma, const float* d_beta,
                  int relu, void* stream);
/* The same pass with the batch statistics taken from the float64 sums the producing convolution's epilogue gathered
 * (ifcb_conv_desc.d_stats: d_sums[c] = sum of z, d_sums[C + c] = sum of z^2 over batch*H*W pixels): mean / invstd are derived
 * on the fly, written to d_mean / d_invstd for the backward pass, and the running statistics are updated as ifcb_bn_stats does
 * (torch.nn.BatchNorm2d train mode).  Replaces ifcb_bn_stats + ifcb_bn_apply: no statistics pass over z. */
int ifcb_bn_apply_sums(const ifcb_view* z, const ifcb_view* out, const ifcb_view* residual, int batch, int dtype,
                       const double* d_sums, float eps, float momentum, float* d_mean, float* d_invstd,
                       float* d_running_mean, float* d_running_var, const float* d_gamma, const float* d_beta, int relu,
                       void* stream);
/* Backward of (BN train -> [+residual] -> [ReLU]) -- autograd's threshold_backward +
 * native_batch_norm_backward:
 *   dy' = dy * [forward output > 0]   (relu = 1; the mask is recomputed from z with the forward's own
 *                                      arithmetic, or read from `a` -- required -- when a residual was added)
 *   dz  = gamma*invstd*(dy' - mean(dy') - xhat*mean(dy'*xhat)),  xhat = (z-mean)*invstd
 *   d_dgamma[C] += sum(dy'*xhat), d_dbeta[C] += sum(dy')
 *   dres (optional, the residual branch's gradient) = or += dy'
 * dz may alias dy.  d_acc: the 64 KB scratch shared with ifcb_bn_stats: float64 [0, 4096) are per-channel
 * accumulators (2*C used; zero on entry, zero again on exit), the 3*C coefficient floats go behind them, and 1 KB at
 * float64 index 7680 holds one arrival counter per block of 64 channels (zero on entry and exit): the reductions run on
 * a (pixel splits x channel blocks) grid and the last block of each channel block finalises its channels in place. */
int ifcb_bn_backward(const ifcb_view* dy, const ifcb_view* a, const ifcb_view* z, const ifcb_view* dz,
                     const ifcb_view* dres, int dres_accumulate, int relu, int batch, int dtype,
                     const float* d_mean, const float* d_invstd, const float* d_gamma, const float* d_beta,
                     double* d_acc, float* d_dgamma, float* d_dbeta, void* stream);
/* The same backward when the BatchNorm read a tensor that OTHER layers read too (DenseNet's pre-activation norm over the
 * growing concatenation, torchvision densenet.py _DenseLayer): dz is ADDED to what the target already holds (dz must not alias dy;
 * no residual route; the ReLU mask is recomputed from z). */
int ifcb_bn_backward_accumulate(const ifcb_view* dy, const ifcb_view* z, const ifcb_view* dz, int relu, int batch, int dtype,
                                const float* d_mean, const float* d_invstd, const float* d_gamma, const float* d_beta,
                                double* d_acc, float* d_dgamma, float* d_dbeta, void* stream);
/* Backward of (Conv2d with bias -> [ReLU]) for the families without BatchNorm (AlexNet, VGG, SqueezeNet; alexnet.py, vgg.py,
 * squeezenet.py under NeustonModel.training_step, neuston_models.py:80-86): dz = dy * [a > 0] (dz may alias dy; relu = 0 copies),
 * d_dbias[C] += sum over pixels of dz.  d_acc: the 64 KB scratch of ifcb_bn_stats. */
int ifcb_bias_relu_backward(const ifcb_view* dy, const ifcb_view* a, const ifcb_view* dz, int relu, int batch, int dtype,
                            double* d_acc, float* d_dbias, void* stream);
/* y = x * scale element-wise; scale float32 [batch, H, W, C] dense (nn.Dropout of the classifier stacks in train mode: the mask
 * times 1/(1-p) from ifcb_dropout_scale; the backward pass applies the same scale to the gradient). */
int ifcb_scale_elems(const ifcb_view* x, const ifcb_view* y, const float* d_scale, int batch, int dtype, void* stream);

/* Pooling for the TRAIN step.  max: F.max_pool2d forward recording the winning tap (first maximum
 * in row-major window order, uint8 [batch,P,Q,C]) and its backward; avg: F.avg_pool2d with
 * count_include_pad=True (divisor k*k) and its backward.  Backward kernels gather (no atomics) and
 * either write dx or add to it (`accumulate`). */
int ifcb_maxpool_fwd_train(const ifcb_view* x, const ifcb_view* y, uint8_t* d_idx, int batch, int k, int stride,
                           int pad, int dtype, void* stream);
int ifcb_maxpool_bwd(const ifcb_view* dy, const uint8_t* d_idx, const ifcb_view* dx, int accumulate, int batch,
                     int k, int stride, int pad, int dtype, void* stream);
int ifcb_avgpool_fwd(const ifcb_view* x, const ifcb_view* y, int batch, int k, int stride, int pad, int dtype,
                     void* stream);
int ifcb_avgpool_bwd(const ifcb_view* dy, const ifcb_view* dx, int accumulate, int batch, int k, int stride,
                     int pad, int dtype, void* stream);

/* out[n, p*stride_h, q*stride_w, :] = in[n,p,q,:] into a zero tensor: the data gradient of a strided
 * conv is the stride-1 conv of the zero-dilated output gradient with the flipped filter, which runs
 * on the forward tcgen05 kernel (ifcb_plan_add_conv with weights from ifcb_conv_repack). */
int ifcb_dilate(const ifcb_view* in, const ifcb_view* out, int batch, int stride_h, int stride_w, void* stream);
/* float32 NCHW [batch,Cin,H,W] -> 16-bit NHWC view (channels >= Cin zero): stem input for ifcb_conv_wgrad. */
int ifcb_nchw_to_nhwc(const float* d_in, int Cin, const ifcb_view* out, int batch, int dtype, void* stream);

/* Patch matrix of the first convolution (Cin = 3): float32 NCHW [batch,3,H,W] -> 16-bit NHWC view
 * [batch, P, Q, K8], channel k = (r*kw + s)*3 + c holds scale[c]*in[n,c,p*stride-pad+r,q*stride-pad+s]+shift[c]
 * (0 outside the image and for k >= kh*kw*3; h_scale/h_shift: torchvision transform_input, NULL = identity).
 * The stem then runs as a 1x1 convolution with Cin = K8 on the tcgen05 forward / weight-gradient kernels. */
int ifcb_stem_im2col(const float* d_in, int H, int W, const ifcb_view* out, int batch, int kh, int kw, int stride,
                     int pad, const float* h_scale, const float* h_shift, int dtype, void* stream);

/* Train-mode head: adaptive_avg_pool2d(1) -> dropout scale -> Linear -> CrossEntropyLoss(mean)
 * (inception.py:147-153 / resnet.py avgpool+fc; NeustonModel.loss neuston_models.py:70-80).
 *   d_dropscale  float32 [batch,C] inverted-dropout scale (ifcb_dropout_scale) or NULL
 *   d_labels     int64 [batch]
 *   loss_weight  1.0 for the main head, 0.4 for Inception's AuxLogits (neuston_models.py:76)
 *   d_pooled [batch,C], d_logits [batch,n_classes] (optional), d_dlogits [batch,n_classes] =
 *   loss_weight/batch * (softmax - onehot); *d_loss += loss_weight * mean CE. */
int ifcb_head_train_fwd(const ifcb_view* x, int batch, int dtype, const float* d_dropscale, const float* d_weight,
                        const float* d_bias, const int64_t* d_labels, int n_classes, float loss_weight,
                        float* d_pooled, float* d_logits, float* d_dlogits, float* d_loss, void* stream);
/* d_dweight[n_classes,C] += dlogits^T pooled; d_dbias += sum_b dlogits; dx (=|+=) W^T dlogits * dropscale / HW */
int ifcb_head_bwd(const ifcb_view* dx, int dx_accumulate, int batch, int dtype, const float* d_dropscale,
                  const float* d_weight, const float* d_pooled, const float* d_dlogits, int n_classes,
                  float* d_dweight, float* d_dbias, void* stream);
int ifcb_dropout_scale(float* d_out, int64_t n, float p, uint64_t seed, void* stream);

/* torch.optim.Adam(lr, betas, eps; no weight decay) over a flat fp32 arena (configure_optimizers,
 * neuston_models.py:63-64); grad_scale multiplies the gradient first (1/world_size after a SUM all-reduce). */
int ifcb_adam_step(float* d_param, const float* d_grad, float* d_m, float* d_v, int64_t n, float lr, float beta1,
                   float beta2, float eps, int step, float grad_scale, void* stream);
/* fp32 master conv weights [Cout, kh*kw, Cin] -> 16-bit tensor-core operands:
 *   d_wfwd    [*, kh*kw*Cin_pad]   forward operand of ifcb_plan_add_conv (NULL to skip)
 *   d_wdgrad  [*, kh*kw*Cout_padk] data-gradient operand: taps reversed, Cin/Cout swapped (NULL to skip)
 * Padding entries are never written (allocate zeroed). */
int ifcb_conv_repack(const float* d_master, int Cout, int taps, int Cin, void* d_wfwd, int Cin_pad, void* d_wdgrad,
                     int Cout_padk, int dtype, void* stream);
/* The same for every conv of a network in one launch: d_items = DEVICE array of n_items records (arguments as
 * ifcb_conv_repack; NULL operands are skipped). */
typedef struct {
  const float* d_master;
  void* d_wfwd;
  void* d_wdgrad;
  int32_t Cout, taps, Cin, Cin_pad, Cout_padk, reserved;
} ifcb_repack_item;
int ifcb_conv_repack_batch(const ifcb_repack_item* d_items, int n_items, int dtype, void* stream);

/* Stem: first convolution (Cin = 3) computed directly in fp32 on CUDA cores from either
 * the resized gray plane (u8) or a float32 NCHW [batch,3,H,W] tensor (drop-in forward(x)).
 * Output 16-bit NHWC.
 *   d_weight float32 [kh*kw*3, Cout] (k = (r*kw+s)*3 + c), d_scale/d_shift [Cout]
 *   f32 input: x_c = in[c] * in_scale[c] + in_shift[c] (torchvision transform_input,
 *              inception.py:95-101; identity otherwise).
 *   u8 input : the three channels are affine in the gray level g,
 *              x_c = a_c * g/255 + b_c (ToTensor, --img-norm, transform_input), so the host
 *              folds the weights to one channel: d_wgray[tap][co] = sum_c w*a_c/255 and
 *              d_wconst[tap][co] = sum_c w*b_c (float64 -> float32).  d_wconst is read only
 *              when pad > 0 (per in-bounds tap); for pad == 0 the host adds
 *              scale*sum_tap wconst to d_shift.  d_weight is unused for u8 input.
 */
enum { IFCB_STEM_IN_U8_GRAY = 0, IFCB_STEM_IN_F32_NCHW = 1 };
typedef struct {
  const void* d_in;
  int32_t in_kind;
  int32_t batch_cap, H, W;
  int32_t kh, kw, stride, pad;
  int32_t Cout;             /* 32, 64 or 96 */
  const float* d_weight;
  const float* d_scale;
  const float* d_shift;
  const float* d_wgray;     /* u8 input only: [kh*kw, Cout] */
  const float* d_wconst;    /* u8 input with pad > 0: [kh*kw, Cout] */
  float in_scale[3], in_shift[3];
  void* d_out;
  int32_t out_ld;
  int32_t relu;
  int32_t dtype; /* IFCB_ACT_* of the output */
  int32_t out_pad_h, out_pad_w; /* zero border of the output tensor (see ifcb_conv_segment) */
} ifcb_stem_desc;
int ifcb_plan_add_stem(ifcb_plan* plan, const ifcb_stem_desc* desc);

/* K3/K4  pooling over bf16 NHWC views.
 *   IFCB_POOL_MAX: max over the window (padding ignored, as torch max_pool2d).
 *   IFCB_POOL_AVG_AFFINE: F.avg_pool2d(count_include_pad=True) (inception.py:204,
 *     278,356) applied AFTER the branch's 1x1 convolution (the two commute), then
 *     the branch's folded BN scale/shift and ReLU: y = relu(scale*avg(x)+shift).
 *     With k = 1 it is a per-channel affine (+ReLU): DenseNet's pre-activation norm -> relu
 *     (densenet.py _DenseLayer / _Transition); with k = 2, stride 2, scale 1 its transition pool.
 */
enum { IFCB_POOL_MAX = 0, IFCB_POOL_AVG_AFFINE = 1 };
typedef struct {
  int32_t kind;
  const void* d_in;
  int32_t in_ld, C;
  int32_t batch_cap, H, W;
  int32_t k, stride, pad;
  void* d_out;
  int32_t out_ld;
  const float* d_scale; /* AVG_AFFINE only */
  const float* d_shift;
  int32_t relu;
  int32_t dtype; /* IFCB_ACT_* of input and output */
  int32_t in_pad_h, in_pad_w, out_pad_h, out_pad_w; /* zero borders of the two tensors */
  int32_t ceil_mode; /* 1: output extent as torch's ceil_mode=True (windows may overhang the bottom / right edge;
                        squeezenet.py MaxPool2d(3, 2, ceil_mode=True)) */
} ifcb_pool_desc;
int ifcb_plan_add_pool(ifcb_plan* plan, const ifcb_pool_desc* desc);

/* K6  head: global average pool -> Linear -> softmax(dim=1) -> top-1
 * (inception.py:147-153 / resnet avgpool+fc; neuston_models.py:155-156;
 * neuston_callbacks.py:161-162).
 *   d_in       bf16 NHWC [batch, HW, C] (pixel stride in_ld)
 *   d_weight   float32 [n_classes, C], d_bias float32 [n_classes]
 *   d_scores   float32 [batch, n_classes] softmax probabilities
 *   d_logits   optional float32 [batch, n_classes] (NULL to skip)
 *   d_top1     int32 [batch] argmax; d_top1_score float32 [batch]
 */
typedef struct {
  const void* d_in;
  int32_t in_ld, C, HW;
  int32_t batch_cap, n_classes;
  const float* d_weight;
  const float* d_bias;
  float* d_scores;
  float* d_logits;
  int32_t* d_top1;
  float* d_top1_score;
  int32_t dtype; /* IFCB_ACT_* of the input */
} ifcb_head_desc;
int ifcb_plan_add_head(ifcb_plan* plan, const ifcb_head_desc* desc);

/* Test-only: issue one im2col TMA load through conv layer `layer`'s tensor map
 * (base pixel (w,h,n), channel c, filter-tap offsets) and copy the raw 128 x 64
 * bf16 shared-memory tile (128B-swizzled) to d_out (16 KiB). */
int ifcb_debug_im2col_probe(ifcb_plan* plan, int layer, int c, int w, int h, int n,
                            int off_w, int off_h, void* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IFCB_B200_H */
