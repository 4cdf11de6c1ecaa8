// Network plan: host-side C ABI (include/ifcb_b200.h) that records layers over
// caller-owned device buffers, encodes their TMA descriptors once, and replays
// them as a fixed launch sequence on a stream (CUDA-graph capturable).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>
#include <memory>
#include <vector>
#include "layers.cuh"

namespace ifcb {

// cuTensorMapEncode* resolved through the runtime so that the library has no
// link-time dependency on libcuda (it must load on a CPU-only build box).
PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;
PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col = nullptr;

int resolve_driver() {
  if (g_encode_tiled && g_encode_im2col) return 0;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  IFCB_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  IFCB_ARG_CHECK(q == cudaDriverEntryPointSuccess && fn, "cuTensorMapEncodeTiled not available");
  g_encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  fn = nullptr;
  IFCB_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  IFCB_ARG_CHECK(q == cudaDriverEntryPointSuccess && fn, "cuTensorMapEncodeIm2col not available");
  g_encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fn);
  return 0;
}

namespace {

enum LayerKind { kConv = 0, kStem = 1, kPool = 2, kHead = 3 };

struct Layer {
  LayerKind kind;
  ConvLayer conv;
  StemLayer stem;
  PoolLayer pool;
  HeadLayer head;
};

inline int out_dim(int in, int k, int stride, int pad) { return (in + 2 * pad - k) / stride + 1; }

int pick_tile_n(int Cout, int hint) {
  if (hint > 0) return hint;
  const int c16 = (Cout + 15) & ~15;
  if (c16 <= 256) return c16;
  // smallest number of equal tiles of width <= 256 (multiple of 16)
  for (int t = 2; t <= 64; ++t) {
    int w = (((c16 + t - 1) / t) + 15) & ~15;
    if (w <= 256) return w;
  }
  return 256;
}

// Heuristic algorithm / N-tile choice (measured on B200, profiles/r01_layer_matrix_v3.txt).
//  * WINDOW pays for every padded anchor row (junk = Hp*Wp / (P*Q)) but reads the input patch
//    once for all taps and lets m accumulators share a weight tile; it wins for >= 4 taps on
//    large images.  1x1 convs and small images (17x17 1x7, 8x8 3x3) stay on IM2COL.
//  * A 128-row tile streams the whole weight matrix at 64 B per MMA clock -- the L2 limit --
//    so WINDOW layers wider than 128 channels are split into N tiles <= 128 to make room in
//    TMEM for two accumulators per weight tile.
//  * IM2COL runs on CTA pairs (256-pixel tiles, N tile <= 256 split over the two SMs).
// WINDOW layers with >= 64 output channels run on CTA pairs (an internal refinement of IFCB_CONV_WINDOW: same requirements on the
// input; each SM holds and fills half of every weight tile, and layers wider than 128 channels keep ONE N tile of up to 256
// channels instead of two).  Measured at batch 1024 (profiles/r02_conv_window_pair_ab.txt): Conv2d_4a 1757 -> 1348 us -- the 20 % the
// shared-memory-port model of tools/conv_rooflines.py predicts --, the 35^2 3x3 / 5x5 layers -12 ... -20 %; the 64-byte-row layers
// (Cin <= 32: Conv2d_2b) lose 17 % and stay on one CTA.  IFCB_CONV_WINDOW_PAIR=0 switches it off.
bool window_pair(int Cout, int Cin) {
  static const bool on = !(getenv("IFCB_CONV_WINDOW_PAIR") && atoi(getenv("IFCB_CONV_WINDOW_PAIR")) == 0);
  return on && (Cout >= 128 || (Cout >= 64 && Cin > 32));
}

int auto_tile_n(int Cout, int algo) {
  const int c16 = (Cout + 15) & ~15;
  if (algo == IFCB_CONV_WINDOW && window_pair(Cout, 64) && Cout >= 128) algo = IFCB_CONV_IM2COL_PAIR;      // one N tile of up to 256 channels
  if (algo == IFCB_CONV_WINDOW && c16 > 128) {
    const int t = (c16 + 127) / 128;
    return (((c16 + t - 1) / t) + 15) & ~15;
  }
  if (algo == IFCB_CONV_IM2COL_PAIR) {
    const int t = (c16 + 255) / 256;
    return (((c16 + t - 1) / t) + 15) & ~15;        // multiple of 16: each SM holds tile_n/2 rows
  }
  return pick_tile_n(Cout, 0);
}

void auto_config(int H, int W, int Cin, int Cout, int kh, int kw, int stride_h, int stride_w, int pad_h, int pad_w,
                 int* algo, int* tile_n) {
  const int P = out_dim(H, kh, stride_h, pad_h), Q = out_dim(W, kw, stride_w, pad_w);
  const int taps = kh * kw;
  const double junk = (double)(H + 2 * pad_h) * (W + 2 * pad_w) / ((double)P * Q);
  const bool window = stride_h == 1 && stride_w == 1 && taps >= 4 && junk <= 1.30;
  static const bool nopair = getenv("IFCB_CONV_NOPAIR") != nullptr;
  // CTA pairs pay off only with a deep K loop and a wide N tile (measured: K >= 512, Cout >= 128)
  const bool pair = !nopair && taps * ((Cin + 63) / 64) >= 8 && Cout >= 128;
  const int a = window ? IFCB_CONV_WINDOW : (pair ? IFCB_CONV_IM2COL_PAIR : IFCB_CONV_IM2COL);
  if (algo) *algo = a;
  if (tile_n) *tile_n = auto_tile_n(Cout, a);
}

}  // namespace
}  // namespace ifcb

struct ifcb_plan {
  std::vector<ifcb::Layer> layers;
};

using namespace ifcb;

extern "C" int ifcb_plan_create(ifcb_plan** out) {
  IFCB_ARG_CHECK(out != nullptr, "ifcb_plan_create: null out");
  *out = new ifcb_plan();
  return 0;
}

extern "C" int ifcb_plan_destroy(ifcb_plan* plan) {
  delete plan;
  return 0;
}

extern "C" int ifcb_plan_num_layers(const ifcb_plan* plan) { return plan ? (int)plan->layers.size() : -1; }
extern "C" int ifcb_plan_num_launches(const ifcb_plan* plan) { return plan ? (int)plan->layers.size() : -1; }

extern "C" int ifcb_conv_geometry(int Cin, int Cout, int kh, int kw, int tile_n_hint, int32_t* Cin_pad,
                                  int32_t* K_pad, int32_t* tile_n, int32_t* Cout_pad) {
  IFCB_ARG_CHECK(Cin > 0 && Cout > 0 && kh > 0 && kw > 0, "ifcb_conv_geometry: bad shape");
  IFCB_ARG_CHECK(tile_n_hint == 0 || (tile_n_hint % 16 == 0 && tile_n_hint >= 16 && tile_n_hint <= 256),
                 "ifcb_conv_geometry: tile_n must be a multiple of 16 in [16,256]");
  const int cp = Cin <= 32 ? 32 : ((Cin + 63) & ~63);      // 64-byte operand rows for the thin first layers
  const int tn = pick_tile_n(Cout, tile_n_hint);
  if (Cin_pad) *Cin_pad = cp;
  if (K_pad) *K_pad = kh * kw * cp;
  if (tile_n) *tile_n = tn;
  if (Cout_pad) *Cout_pad = ((Cout + tn - 1) / tn) * tn;
  return 0;
}

extern "C" int ifcb_conv_auto_config(int H, int W, int Cin, int Cout, int kh, int kw, int stride_h, int stride_w, int pad_h,
                                     int pad_w, int32_t* algo, int32_t* tile_n) {
  IFCB_ARG_CHECK(H > 0 && W > 0 && Cin > 0 && Cout > 0 && kh > 0 && kw > 0 && stride_h > 0 && stride_w > 0 && pad_h >= 0 && pad_w >= 0,
                 "ifcb_conv_auto_config: bad shape");
  IFCB_ARG_CHECK(out_dim(H, kh, stride_h, pad_h) > 0 && out_dim(W, kw, stride_w, pad_w) > 0,
                 "ifcb_conv_auto_config: empty output");
  int a = 0, t = 0;
  auto_config(H, W, Cin, Cout, kh, kw, stride_h, stride_w, pad_h, pad_w, &a, &t);
  if (algo) *algo = a;
  if (tile_n) *tile_n = t;
  return 0;
}

extern "C" int ifcb_conv_auto_tile_n(int Cout, int algo) {
  if (Cout <= 0 || algo < IFCB_CONV_IM2COL || algo > IFCB_CONV_IM2COL_PAIR) return -1;
  return ifcb::auto_tile_n(Cout, algo);
}

extern "C" int ifcb_plan_add_conv(ifcb_plan* plan, const ifcb_conv_desc* d) {
  IFCB_ARG_CHECK(plan && d, "ifcb_plan_add_conv: null argument");
  IFCB_ARG_CHECK(d->d_in && d->d_weight && d->d_scale && d->d_shift, "conv: null tensor pointer");
  IFCB_ARG_CHECK(d->Cin > 0 && d->Cin % 8 == 0, "conv: Cin=%d must be a positive multiple of 8", d->Cin);
  IFCB_ARG_CHECK(d->in_ld >= d->Cin && d->in_ld % 8 == 0, "conv: in_ld=%d must be >= Cin and a multiple of 8",
                 d->in_ld);
  IFCB_ARG_CHECK((reinterpret_cast<uintptr_t>(d->d_in) & 15) == 0, "conv: d_in must be 16-byte aligned");
  IFCB_ARG_CHECK(d->batch_cap > 0 && d->H > 0 && d->W > 0, "conv: bad input shape");
  IFCB_ARG_CHECK(d->kh >= 1 && d->kw >= 1 && d->kh <= 16 && d->kw <= 16, "conv: bad filter %dx%d", d->kh, d->kw);
  IFCB_ARG_CHECK(d->stride_h >= 1 && d->stride_w >= 1 && d->stride_h <= 8 && d->stride_w <= 8, "conv: bad stride");
  IFCB_ARG_CHECK(d->pad_h >= 0 && d->pad_w >= 0 && d->pad_h < d->kh && d->pad_w < d->kw, "conv: bad padding");
  IFCB_ARG_CHECK(d->in_pad_h >= 0 && d->in_pad_w >= 0 && d->in_pad_h <= 8 && d->in_pad_w <= 8, "conv: bad in_pad");
  IFCB_ARG_CHECK(d->n_seg >= 1 && d->n_seg <= IFCB_MAX_SEGMENTS, "conv: n_seg=%d out of range", d->n_seg);
  IFCB_ARG_CHECK(d->tile_n == 0 || (d->tile_n % 16 == 0 && d->tile_n >= 16 && d->tile_n <= 256),
                 "conv: tile_n=%d must be a multiple of 16 in [16,256]", d->tile_n);
  IFCB_ARG_CHECK(d->dtype == IFCB_ACT_BF16 || d->dtype == IFCB_ACT_FP16, "conv: bad dtype %d", d->dtype);
  IFCB_ARG_CHECK(d->algo >= IFCB_CONV_AUTO && d->algo <= IFCB_CONV_IM2COL_PAIR, "conv: bad algo %d", d->algo);
  int rc = resolve_driver();
  if (rc) return rc;

  const bool can_window = d->stride_h == 1 && d->stride_w == 1 && d->in_pad_h >= d->pad_h && d->in_pad_w >= d->pad_w;
  IFCB_ARG_CHECK(d->algo != IFCB_CONV_WINDOW || can_window,
                 "conv: the window algorithm needs stride 1 and an input buffer padded by at least the conv padding");
  int auto_algo = IFCB_CONV_IM2COL, auto_tile_n = 0;
  auto_config(d->H, d->W, d->Cin, d->Cout, d->kh, d->kw, d->stride_h, d->stride_w, d->pad_h, d->pad_w, &auto_algo, &auto_tile_n);
  const bool window = d->algo == IFCB_CONV_WINDOW || (d->algo == IFCB_CONV_AUTO && can_window && auto_algo == IFCB_CONV_WINDOW);
  const bool wpair = window && window_pair(d->Cout, d->Cin);
  const bool pair = wpair || (!window && (d->algo == IFCB_CONV_IM2COL_PAIR || (d->algo == IFCB_CONV_AUTO && auto_algo == IFCB_CONV_IM2COL_PAIR)));
  const int tile_n_req = d->tile_n ? d->tile_n
                                   : ifcb::auto_tile_n(d->Cout, window ? IFCB_CONV_WINDOW : pair ? IFCB_CONV_IM2COL_PAIR : IFCB_CONV_IM2COL);

  Layer L{};
  L.kind = kConv;
  L.conv.window = window;
  L.conv.pair = pair;
  ConvKernelParams& kp = L.conv.kp;
  const int P = out_dim(d->H, d->kh, d->stride_h, d->pad_h);
  const int Q = out_dim(d->W, d->kw, d->stride_w, d->pad_w);
  IFCB_ARG_CHECK(P > 0 && Q > 0, "conv: empty output");
  const int Hp = d->H + 2 * d->in_pad_h, Wp = d->W + 2 * d->in_pad_w;     // physical input extent
  int32_t cin_pad, k_pad, tile_n, cout_pad;
  ifcb_conv_geometry(d->Cin, d->Cout, d->kh, d->kw, tile_n_req, &cin_pad, &k_pad, &tile_n, &cout_pad);
  kp.rows = 0;
  kp.P = P;
  kp.Q = Q;
  kp.rows_per_img = window ? Hp * Wp : P * Q;
  kp.row_w = window ? Wp : Q;
  kp.kh = d->kh; kp.kw = d->kw;
  kp.stride_h = d->stride_h; kp.stride_w = d->stride_w;
  kp.pad_h = d->pad_h; kp.pad_w = d->pad_w;
  const int row_elems = cin_pad == 32 ? 32 : 64;
  kp.row_bytes = row_elems * 2;
  kp.cblocks = cin_pad / row_elems;
  const int last = d->Cin - (kp.cblocks - 1) * row_elems;
  kp.last_ksteps = (last + 15) / 16;
  kp.tile_n = tile_n;
  kp.n_tiles = cout_pad / tile_n;
  kp.cout_pad = cout_pad;
  kp.fp16 = d->dtype;
  {
    const char* e = getenv("IFCB_CONV_DEBUG");
    kp.debug_flags = e ? atoi(e) : 0;
    const char* ms = getenv("IFCB_CONV_MSUB");
    kp.m_sub_cap = ms ? atoi(ms) : 4;
    const char* as = getenv("IFCB_CONV_ASLOTS");
    kp.a_slots_pref = as ? atoi(as) : 2;
    const char* bg = getenv("IFCB_CONV_BGROUP");
    kp.b_group_cap = bg ? atoi(bg) : 0;
  }
  kp.stats = d->d_stats;
  kp.cout = d->Cout;
  kp.n_major = (d->d_stats && kp.n_tiles > 1) ? 1 : 0;
  IFCB_ARG_CHECK(!d->d_stats || d->n_seg == 1, "conv: d_stats needs a single output segment");
  kp.win_shift0 = (d->in_pad_h - d->pad_h) * Wp + (d->in_pad_w - d->pad_w);
  const int halo = kp.win_shift0 + (d->kh - 1) * Wp + (d->kw - 1);
  IFCB_ARG_CHECK(conv_plan_smem(kp, window, pair, halo), "conv: no shared-memory plan for tile_n=%d halo=%d", tile_n, halo);
  kp.scale = d->d_scale;
  kp.shift = d->d_shift;
  kp.residual = reinterpret_cast<const __nv_bfloat16*>(d->d_residual);
  kp.res_ld = d->res_ld;
  kp.res_pad_h = d->res_pad_h;
  kp.res_pad_w = d->res_pad_w;
  if (d->d_residual) {
    IFCB_ARG_CHECK(d->res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(d->d_residual) & 15) == 0,
                   "conv: residual view must be 16-byte aligned with ld %% 8 == 0");
    IFCB_ARG_CHECK(d->n_seg == 1 && d->seg[0].n_begin == 0, "conv: residual needs a single segment at column 0");
    IFCB_ARG_CHECK((unsigned long long)d->batch_cap * (P + 2 * d->res_pad_h) * (Q + 2 * d->res_pad_w) * (unsigned long long)d->res_ld < (1ull << 32) - 1,
                   "conv: residual exceeds the 32-bit element offsets of the epilogue");
  }
  kp.n_seg = d->n_seg;
  for (int s = 0; s < d->n_seg; ++s) {
    const ifcb_conv_segment& sg = d->seg[s];
    IFCB_ARG_CHECK(sg.n_begin % 16 == 0 && sg.n_end % 16 == 0 && sg.n_begin < sg.n_end && sg.n_end <= cout_pad,
                   "conv: segment %d [%d,%d) must be 16-aligned and inside [0,%d)", s, sg.n_begin, sg.n_end, cout_pad);
    IFCB_ARG_CHECK(sg.d_out && sg.ld % 8 == 0 && (reinterpret_cast<uintptr_t>(sg.d_out) & 15) == 0,
                   "conv: segment %d output must be 16-byte aligned with ld %% 8 == 0", s);
    IFCB_ARG_CHECK(sg.ld >= sg.n_end - sg.n_begin, "conv: segment %d ld too small", s);
    IFCB_ARG_CHECK(sg.pad_h >= 0 && sg.pad_w >= 0 && sg.pad_h <= 8 && sg.pad_w <= 8, "conv: segment %d bad pad", s);
    IFCB_ARG_CHECK((unsigned long long)d->batch_cap * (P + 2 * sg.pad_h) * (Q + 2 * sg.pad_w) * (unsigned long long)sg.ld < (1ull << 32) - 1,
                   "conv: segment %d destination exceeds the 32-bit element offsets of the epilogue", s);
    kp.seg_begin[s] = sg.n_begin;
    kp.seg_end[s] = sg.n_end;
    kp.seg_ld[s] = sg.ld;
    kp.seg_relu[s] = sg.relu;
    kp.seg_pad_h[s] = sg.pad_h;
    kp.seg_pad_w[s] = sg.pad_w;
    kp.seg_out[s] = reinterpret_cast<__nv_bfloat16*>(sg.d_out);
  }
  {
    kp.magic_img = div_magic(kp.rows_per_img);
    kp.magic_w = div_magic(kp.row_w);
    bool ident = !window && (!d->d_residual || (d->res_pad_h == 0 && d->res_pad_w == 0));
    for (int sgi = 0; sgi < d->n_seg; ++sgi) ident = ident && d->seg[sgi].pad_h == 0 && d->seg[sgi].pad_w == 0;
    kp.identity_rows = ident ? 1 : 0;
  }
  L.conv.batch_cap = d->batch_cap;
  const CUtensorMapDataType dt = d->dtype == IFCB_ACT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle swz = kp.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;

  if (window) {
    // --- A: tiled 2-D map over the padded input viewed as [batch_cap*Hp*Wp, Cin] ---
    cuuint64_t gdim[2] = {(cuuint64_t)d->Cin, (cuuint64_t)d->batch_cap * Hp * Wp};
    cuuint64_t gstr[1] = {(cuuint64_t)d->in_ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)row_elems, (cuuint32_t)kp.box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&L.conv.tmap_a, dt, 2, const_cast<void*>(d->d_in), gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IFCB_ARG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for window input rows=%d", (int)r, kp.box_rows);
  } else {
    // --- A: im2col tensor map over the interior of the NHWC input (dims C, W, H, N) ---
    const char* base = reinterpret_cast<const char*>(d->d_in) + ((size_t)d->in_pad_h * Wp + d->in_pad_w) * d->in_ld * 2;
    cuuint64_t gdim[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->batch_cap};
    cuuint64_t gstr[3] = {(cuuint64_t)d->in_ld * 2, (cuuint64_t)Wp * d->in_ld * 2, (cuuint64_t)Hp * Wp * d->in_ld * 2};
    int lower[2] = {-d->pad_w, -d->pad_h};
    int upper[2] = {d->pad_w - (d->kw - 1), d->pad_h - (d->kh - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)d->stride_w, (cuuint32_t)d->stride_h, 1};
    CUresult r = g_encode_im2col(&L.conv.tmap_a, dt, 4, const_cast<char*>(base), gdim, gstr, lower, upper,
                                 /*channelsPerPixel=*/(cuuint32_t)row_elems, /*pixelsPerColumn=*/128, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IFCB_ARG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (%d) for conv %dx%d Cin=%d H=%d W=%d", (int)r,
                   d->kh, d->kw, d->Cin, d->H, d->W);
    // Small-tensor workaround used by CUTLASS for drivers <= 13.1: tensors under
    // 128 KiB must not have bit 21 of the second descriptor word set.
    const unsigned long long bytes = (unsigned long long)d->batch_cap * Hp * Wp * d->in_ld * 2ull;
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && bytes < 131072ull) reinterpret_cast<uint64_t*>(&L.conv.tmap_a)[1] &= ~(1ull << 21);
  }
  // --- B: tiled tensor map over packed weights [Cout_pad, K_pad] ---
  {
    cuuint64_t gdim[2] = {(cuuint64_t)k_pad, (cuuint64_t)cout_pad};
    cuuint64_t gstr[1] = {(cuuint64_t)k_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)row_elems, (cuuint32_t)(pair ? tile_n / 2 : tile_n)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&L.conv.tmap_b, dt, 2, const_cast<void*>(d->d_weight), gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IFCB_ARG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for weights K_pad=%d Cout_pad=%d", (int)r,
                   k_pad, cout_pad);
  }
  plan->layers.push_back(L);
  return 0;
}

extern "C" int ifcb_plan_add_stem(ifcb_plan* plan, const ifcb_stem_desc* d) {
  IFCB_ARG_CHECK(plan && d, "ifcb_plan_add_stem: null argument");
  IFCB_ARG_CHECK(d->d_in && d->d_scale && d->d_shift && d->d_out, "stem: null tensor pointer");
  IFCB_ARG_CHECK(d->Cout == 32 || d->Cout == 64 || d->Cout == 96, "stem: Cout=%d unsupported (32, 64 or 96)", d->Cout);
  IFCB_ARG_CHECK(d->in_kind == IFCB_STEM_IN_U8_GRAY || d->in_kind == IFCB_STEM_IN_F32_NCHW, "stem: bad in_kind");
  IFCB_ARG_CHECK(d->in_kind != IFCB_STEM_IN_U8_GRAY || (d->d_wgray && (d->pad == 0 || d->d_wconst)),
                 "stem: u8 input needs d_wgray (and d_wconst when pad > 0)");
  IFCB_ARG_CHECK(d->in_kind != IFCB_STEM_IN_F32_NCHW || d->d_weight, "stem: f32 input needs d_weight");
  IFCB_ARG_CHECK(d->out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(d->d_out) & 15) == 0,
                 "stem: output must be 16-byte aligned with ld %% 8 == 0");
  Layer L{};
  L.kind = kStem;
  L.stem.d = *d;
  L.stem.P = out_dim(d->H, d->kh, d->stride, d->pad);
  L.stem.Q = out_dim(d->W, d->kw, d->stride, d->pad);
  IFCB_ARG_CHECK(L.stem.P > 0 && L.stem.Q > 0, "stem: empty output");
  if (d->in_kind == IFCB_STEM_IN_U8_GRAY && d->kh == 3 && d->kw == 3 && d->pad == 0) {
    // fast path: the folded gray weights + BN affine become a kernel parameter (constant bank);
    // one synchronous device->host copy at plan-build time
    const int co = d->Cout;
    L.stem.h_const.resize(9 * co + 2 * co);
    IFCB_CUDA_CHECK(cudaMemcpy(L.stem.h_const.data(), d->d_wgray, sizeof(float) * 9 * co, cudaMemcpyDeviceToHost));
    IFCB_CUDA_CHECK(cudaMemcpy(L.stem.h_const.data() + 9 * co, d->d_scale, sizeof(float) * co, cudaMemcpyDeviceToHost));
    IFCB_CUDA_CHECK(cudaMemcpy(L.stem.h_const.data() + 10 * co, d->d_shift, sizeof(float) * co, cudaMemcpyDeviceToHost));
  }
  plan->layers.push_back(L);
  return 0;
}

// After the caller has rewritten weights / folded BN vectors IN PLACE (same device buffers): re-reads the few values the plan
// keeps on the host (the constant-bank stem's folded gray weights).  Synchronous; the caller has synchronised its stream.
extern "C" int ifcb_plan_refresh(ifcb_plan* plan) {
  IFCB_ARG_CHECK(plan != nullptr, "ifcb_plan_refresh: null plan");
  for (auto& L : plan->layers) {
    if (L.kind != kStem || L.stem.h_const.empty()) continue;
    const ifcb_stem_desc& d = L.stem.d;
    const int co = d.Cout;
    IFCB_CUDA_CHECK(cudaMemcpy(L.stem.h_const.data(), d.d_wgray, sizeof(float) * 9 * co, cudaMemcpyDeviceToHost));
    IFCB_CUDA_CHECK(cudaMemcpy(L.stem.h_const.data() + 9 * co, d.d_scale, sizeof(float) * co, cudaMemcpyDeviceToHost));
    IFCB_CUDA_CHECK(cudaMemcpy(L.stem.h_const.data() + 10 * co, d.d_shift, sizeof(float) * co, cudaMemcpyDeviceToHost));
  }
  return 0;
}

extern "C" int ifcb_plan_add_pool(ifcb_plan* plan, const ifcb_pool_desc* d) {
  IFCB_ARG_CHECK(plan && d, "ifcb_plan_add_pool: null argument");
  IFCB_ARG_CHECK(d->d_in && d->d_out, "pool: null tensor pointer");
  IFCB_ARG_CHECK(d->kind == IFCB_POOL_MAX || d->kind == IFCB_POOL_AVG_AFFINE, "pool: bad kind %d", d->kind);
  IFCB_ARG_CHECK(d->kind == IFCB_POOL_MAX || (d->d_scale && d->d_shift), "pool: avg needs scale/shift");
  IFCB_ARG_CHECK(d->C > 0 && d->C % 8 == 0 && d->in_ld % 8 == 0 && d->out_ld % 8 == 0,
                 "pool: C, in_ld, out_ld must be multiples of 8");
  IFCB_ARG_CHECK(((reinterpret_cast<uintptr_t>(d->d_in) | reinterpret_cast<uintptr_t>(d->d_out)) & 15) == 0,
                 "pool: views must be 16-byte aligned");
  IFCB_ARG_CHECK(d->in_pad_h >= 0 && d->in_pad_w >= 0 && d->out_pad_h >= 0 && d->out_pad_w >= 0 &&
                 d->in_pad_h <= 8 && d->in_pad_w <= 8 && d->out_pad_h <= 8 && d->out_pad_w <= 8, "pool: bad pads");
  Layer L{};
  L.kind = kPool;
  L.pool.d = *d;
  L.pool.P = out_dim(d->H, d->k, d->stride, d->pad);
  L.pool.Q = out_dim(d->W, d->k, d->stride, d->pad);
  if (d->ceil_mode) {      // torch pooling_output_shape: ceil division, and the last window must start inside the (left-padded) input
    auto ceil_dim = [](int in, int k, int s, int p) {
      int o = (in + 2 * p - k + s - 1) / s + 1;
      if ((o - 1) * s >= in + p) --o;
      return o;
    };
    L.pool.P = ceil_dim(d->H, d->k, d->stride, d->pad);
    L.pool.Q = ceil_dim(d->W, d->k, d->stride, d->pad);
  }
  IFCB_ARG_CHECK(L.pool.P > 0 && L.pool.Q > 0, "pool: empty output");
  plan->layers.push_back(L);
  return 0;
}

extern "C" int ifcb_plan_add_head(ifcb_plan* plan, const ifcb_head_desc* d) {
  IFCB_ARG_CHECK(plan && d, "ifcb_plan_add_head: null argument");
  IFCB_ARG_CHECK(d->d_in && d->d_weight && d->d_bias && d->d_scores && d->d_top1 && d->d_top1_score,
                 "head: null tensor pointer");
  IFCB_ARG_CHECK(d->C > 0 && d->C % 8 == 0 && d->in_ld % 8 == 0, "head: C and in_ld must be multiples of 8");
  IFCB_ARG_CHECK(d->n_classes > 0 && (d->C + d->n_classes) * 4 <= 200 * 1024, "head: n_classes out of range");
  IFCB_ARG_CHECK((reinterpret_cast<uintptr_t>(d->d_weight) & 15) == 0, "head: weight must be 16-byte aligned");
  Layer L{};
  L.kind = kHead;
  L.head.d = *d;
  plan->layers.push_back(L);
  return 0;
}

static int plan_run_impl(ifcb_plan* plan, int first, int last, int batch, int out_row, void* stream_v);

extern "C" int ifcb_plan_run_range(ifcb_plan* plan, int first, int last, int batch, void* stream_v) {
  return plan_run_impl(plan, first, last, batch, 0, stream_v);
}

extern "C" int ifcb_plan_run_at(ifcb_plan* plan, int batch, int out_row, void* stream_v) {
  IFCB_ARG_CHECK(plan != nullptr, "ifcb_plan_run_at: null plan");
  IFCB_ARG_CHECK(out_row >= 0, "ifcb_plan_run_at: out_row < 0");
  return plan_run_impl(plan, 0, (int)plan->layers.size(), batch, out_row, stream_v);
}

static int plan_run_impl(ifcb_plan* plan, int first, int last, int batch, int out_row, void* stream_v) {
  IFCB_ARG_CHECK(plan != nullptr, "ifcb_plan_run: null plan");
  IFCB_ARG_CHECK(first >= 0 && last <= (int)plan->layers.size() && first <= last, "ifcb_plan_run: bad layer range");
  IFCB_ARG_CHECK(batch >= 0, "ifcb_plan_run: batch < 0");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  for (int i = first; i < last; ++i) {
    const Layer& L = plan->layers[i];
    int rc = 0;
    switch (L.kind) {
      case kConv:
        IFCB_ARG_CHECK(batch <= L.conv.batch_cap, "layer %d: batch %d exceeds capacity %d", i, batch, L.conv.batch_cap);
        rc = launch_conv(L.conv, batch, stream);
        break;
      case kStem:
        IFCB_ARG_CHECK(batch <= L.stem.d.batch_cap, "layer %d: batch %d exceeds capacity", i, batch);
        rc = launch_stem(L.stem, batch, stream);
        break;
      case kPool:
        IFCB_ARG_CHECK(batch <= L.pool.d.batch_cap, "layer %d: batch %d exceeds capacity", i, batch);
        rc = launch_pool(L.pool, batch, stream);
        break;
      case kHead:
        IFCB_ARG_CHECK(batch <= L.head.d.batch_cap, "layer %d: batch %d exceeds capacity", i, batch);
        rc = launch_head(L.head, batch, out_row, stream);
        break;
    }
    if (rc) return rc;
  }
  return 0;
}

extern "C" int ifcb_plan_run(ifcb_plan* plan, int batch, void* stream) {
  IFCB_ARG_CHECK(plan != nullptr, "ifcb_plan_run: null plan");
  return ifcb_plan_run_range(plan, 0, (int)plan->layers.size(), batch, stream);
}

// Test-only probe: one im2col TMA load through conv layer `layer`'s tensor map;
// d_out receives the raw (swizzled) 128x64 bf16 tile.
extern "C" int ifcb_debug_im2col_probe(ifcb_plan* plan, int layer, int c, int w, int h, int n, int off_w, int off_h,
                                       void* d_out, void* stream) {
  IFCB_ARG_CHECK(plan && layer >= 0 && layer < (int)plan->layers.size(), "probe: bad layer");
  IFCB_ARG_CHECK(plan->layers[layer].kind == kConv, "probe: layer %d is not a conv", layer);
  return launch_im2col_probe(plan->layers[layer].conv, c, w, h, n, off_w, off_h, d_out,
                             reinterpret_cast<cudaStream_t>(stream));
}
