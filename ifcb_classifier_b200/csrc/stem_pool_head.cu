// Stem convolution (Cin = 3, direct fp32), pooling (K3/K4) and the fused head
// (K6: global average pool -> Linear -> softmax -> top-1).  All HBM-bound
// integer/elementwise work: coalesced 16-byte accesses over NHWC bf16, warp
// shuffles for the reductions.  See include/ifcb_b200.h for the reference code
// each replaces.
#include "layers.cuh"
#include "ptx.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace ifcb {
namespace {

// ------------------------------------------------------------------------------
// Stem: direct convolution of the 3-channel network input.
//   u8 gray input: x_c(pixel) = lut[c][g] where lut holds ToTensor + Normalize
//   (+ torchvision transform_input) evaluated exactly as torch does in fp32.
//   One thread = one output pixel x COUT channels (accumulators in registers).
// ------------------------------------------------------------------------------
template <int COUT, bool U8>
__global__ void __launch_bounds__(128) stem_kernel(const ifcb_stem_desc d, const float* __restrict__ lut_g,
                                                   int P, int Q, long long total) {
  extern __shared__ float sm[];
  const int taps = d.kh * d.kw;
  float* w_s = sm;                       // [taps*3][COUT]
  float* lut = sm + taps * 3 * COUT;     // [3][256] (U8 only)
  for (int i = threadIdx.x; i < taps * 3 * COUT; i += blockDim.x) w_s[i] = d.d_weight[i];
  if (U8)
    for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = lut_g[i];
  __syncthreads();

  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int PQ = P * Q;
  const int img = (int)(pix / PQ);
  const int rem = (int)(pix - (long long)img * PQ);
  const int op = rem / Q, oq = rem - op * Q;
  const int h0 = op * d.stride - d.pad, w0 = oq * d.stride - d.pad;

  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;

  const uint8_t* in_u8 = reinterpret_cast<const uint8_t*>(d.d_in) + (long long)img * d.H * d.W;
  const float* in_f = reinterpret_cast<const float*>(d.d_in) + (long long)img * 3 * d.H * d.W;
  const long long plane = (long long)d.H * d.W;

  for (int r = 0; r < d.kh; ++r) {
    const int hh = h0 + r;
    if (hh < 0 || hh >= d.H) continue;
    for (int s = 0; s < d.kw; ++s) {
      const int ww = w0 + s;
      if (ww < 0 || ww >= d.W) continue;
      float x0, x1, x2;
      if (U8) {
        const int g = in_u8[(long long)hh * d.W + ww];
        x0 = lut[g];
        x1 = lut[256 + g];
        x2 = lut[512 + g];
      } else {
        const long long o = (long long)hh * d.W + ww;
        x0 = fmaf(in_f[o], d.in_scale[0], d.in_shift[0]);
        x1 = fmaf(in_f[plane + o], d.in_scale[1], d.in_shift[1]);
        x2 = fmaf(in_f[2 * plane + o], d.in_scale[2], d.in_shift[2]);
      }
      const float* wk = w_s + (r * d.kw + s) * 3 * COUT;
#pragma unroll
      for (int c = 0; c < COUT; ++c)
        acc[c] = fmaf(x2, wk[2 * COUT + c], fmaf(x1, wk[COUT + c], fmaf(x0, wk[c], acc[c])));
    }
  }
  const long long orow = ((long long)img * (P + 2 * d.out_pad_h) + op + d.out_pad_h) * (Q + 2 * d.out_pad_w) + oq + d.out_pad_w;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.d_out) + orow * d.out_ld;
#pragma unroll
  for (int c = 0; c < COUT; c += 8) {
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      y[j] = fmaf(acc[c + j], d.d_scale[c + j], d.d_shift[c + j]);
      if (d.relu) y[j] = fmaxf(y[j], 0.f);
    }
    uint4 o;
    o.x = pack_act2(y[0], y[1], d.dtype);
    o.y = pack_act2(y[2], y[3], d.dtype);
    o.z = pack_act2(y[4], y[5], d.dtype);
    o.w = pack_act2(y[6], y[7], d.dtype);
    *reinterpret_cast<uint4*>(out + c) = o;
  }
}

// ------------------------------------------------------------------------------
// Stem, gray fast path.  The three network input channels are affine functions of
// the SAME gray level g: x_c = a_c * g/255 + b_c (ToTensor, --img-norm, transform_input).
// So sum_c w[co][c][tap] * x_c = wg[tap][co] * g + wc[tap][co] with host-folded
//   wg = sum_c w*a_c/255,  wc = sum_c w*b_c   (computed in float64):
// a 1-channel convolution with kh*kw FMAs per output instead of 3*kh*kw.  For pad == 0
// every tap is in bounds and sum_tap wc is folded into the BN shift by the host
// (HASPAD = false); with padding the constant is added per in-bounds tap.
// ------------------------------------------------------------------------------
template <int COUT, bool HASPAD>
__global__ void __launch_bounds__(128) stem_gray_kernel(const ifcb_stem_desc d, int P, int Q, long long total) {
  extern __shared__ float sm[];
  const int taps = d.kh * d.kw;
  float* wg = sm;                        // [taps][COUT]
  float* wc = sm + taps * COUT;          // [taps][COUT] (HASPAD only)
  for (int i = threadIdx.x; i < taps * COUT; i += blockDim.x) {
    wg[i] = d.d_wgray[i];
    if (HASPAD) wc[i] = d.d_wconst[i];
  }
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int PQ = P * Q;
  const int img = (int)(pix / PQ);
  const int rem = (int)(pix - (long long)img * PQ);
  const int op = rem / Q, oq = rem - op * Q;
  const int h0 = op * d.stride - d.pad, w0 = oq * d.stride - d.pad;
  const uint8_t* in = reinterpret_cast<const uint8_t*>(d.d_in) + (long long)img * d.H * d.W;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
  for (int r = 0; r < d.kh; ++r) {
    const int hh = h0 + r;
    if (HASPAD && (hh < 0 || hh >= d.H)) continue;
    for (int s = 0; s < d.kw; ++s) {
      const int ww = w0 + s;
      if (HASPAD && (ww < 0 || ww >= d.W)) continue;
      const float g = (float)in[(long long)hh * d.W + ww];
      const float4* wk = reinterpret_cast<const float4*>(wg + (r * d.kw + s) * COUT);
      const float4* ck = reinterpret_cast<const float4*>(wc + (r * d.kw + s) * COUT);
#pragma unroll
      for (int c4 = 0; c4 < COUT / 4; ++c4) {
        const float4 w = wk[c4];
        if (HASPAD) {
          const float4 k = ck[c4];
          acc[4 * c4 + 0] += fmaf(g, w.x, k.x);
          acc[4 * c4 + 1] += fmaf(g, w.y, k.y);
          acc[4 * c4 + 2] += fmaf(g, w.z, k.z);
          acc[4 * c4 + 3] += fmaf(g, w.w, k.w);
        } else {
          acc[4 * c4 + 0] = fmaf(g, w.x, acc[4 * c4 + 0]);
          acc[4 * c4 + 1] = fmaf(g, w.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(g, w.z, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(g, w.w, acc[4 * c4 + 3]);
        }
      }
    }
  }
  const long long orow = ((long long)img * (P + 2 * d.out_pad_h) + op + d.out_pad_h) * (Q + 2 * d.out_pad_w) + oq + d.out_pad_w;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.d_out) + orow * d.out_ld;
#pragma unroll
  for (int c = 0; c < COUT; c += 8) {
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      y[j] = fmaf(acc[c + j], __ldg(d.d_scale + c + j), __ldg(d.d_shift + c + j));
      if (d.relu) y[j] = fmaxf(y[j], 0.f);
    }
    uint4 o;
    o.x = pack_act2(y[0], y[1], d.dtype);
    o.y = pack_act2(y[2], y[3], d.dtype);
    o.z = pack_act2(y[4], y[5], d.dtype);
    o.w = pack_act2(y[6], y[7], d.dtype);
    *reinterpret_cast<uint4*>(out + c) = o;
  }
}

// ------------------------------------------------------------------------------
// Stem, gray 3x3 / pad 0 (Inception Conv2d_1a): one thread = one output pixel x COUT
// channels.  The 9 x COUT folded gray weights and the BN scale/shift travel as a kernel
// PARAMETER (constant bank): every FFMA reads its weight as a constant operand, so the inner
// loop is 9 byte loads + 9*COUT FFMA with no shared-memory or register traffic for weights
// (FMA-pipe bound).
// ------------------------------------------------------------------------------
template <int COUT>
struct StemConst {
  float w[9 * COUT];
  float scale[COUT], shift[COUT];
};

// d = a * b + d on two packed fp32 lanes (sm_100 FFMA2): one instruction, two multiply-adds
__device__ __forceinline__ void ffma2_acc(float& d0, float& d1, float a, float b0, float b1) {
  asm("{\n.reg .b64 a, b, c;\nmov.b64 a, {%2, %2};\nmov.b64 b, {%3, %4};\nmov.b64 c, {%0, %1};\n"
      "fma.rn.f32x2 c, a, b, c;\nmov.b64 {%0, %1}, c;\n}"
      : "+f"(d0), "+f"(d1)
      : "f"(a), "f"(b0), "f"(b1));
}

template <int COUT>
__global__ void __launch_bounds__(128) stem_gray3x3_kernel(const ifcb_stem_desc d, const __grid_constant__ StemConst<COUT> k,
                                                           int P, int Q, int q2n, int total, unsigned long long magic_pq2,
                                                           unsigned long long magic_q2) {
  // one thread = TWO horizontally adjacent output pixels: every constant-bank weight feeds two FFMAs
  const int idx = blockIdx.x * 128 + threadIdx.x;
  if (idx >= total) return;
  const int img = (int)fast_div((uint32_t)idx, magic_pq2);
  const int rem = idx - img * (P * q2n);
  const int op = (int)fast_div((uint32_t)rem, magic_q2);
  const int oq = (rem - op * q2n) * 2;
  const bool two = oq + 1 < Q;
  const int ncol = 3 + (two ? d.stride : 0);           // input columns this thread touches (<= 5 for stride 2)
  const uint8_t* in = reinterpret_cast<const uint8_t*>(d.d_in) + (long long)img * d.H * d.W + (op * d.stride) * d.W + oq * d.stride;
  float g[3][5];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c) g[r][c] = c < ncol ? (float)__ldg(in + r * d.W + c) : 0.f;
  const int st = d.stride;                             // 1 or 2 (checked by the launcher)
  float acc0[COUT], acc1[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) { acc0[c] = 0.f; acc1[c] = 0.f; }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s2 = 0; s2 < 3; ++s2) {
      const float ga = g[r][s2];
      const float gb = st == 2 ? g[r][s2 + 2] : g[r][s2 + 1];
      // FFMA2: a weight PAIR feeds two packed multiply-adds per pixel -- half the FMA-pipe instructions (measured -5 %)
#pragma unroll
      for (int c = 0; c < COUT; c += 2) {
        const float w0 = k.w[(r * 3 + s2) * COUT + c], w1 = k.w[(r * 3 + s2) * COUT + c + 1];
        ffma2_acc(acc0[c], acc0[c + 1], ga, w0, w1);
        ffma2_acc(acc1[c], acc1[c + 1], gb, w0, w1);
      }
    }
  const long long orow = ((long long)img * (P + 2 * d.out_pad_h) + op + d.out_pad_h) * (Q + 2 * d.out_pad_w) + oq + d.out_pad_w;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.d_out) + orow * d.out_ld;
#pragma unroll
  for (int px = 0; px < 2; ++px) {
    if (px == 1 && !two) break;
#pragma unroll
    for (int c = 0; c < COUT; c += 8) {
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y[j] = fmaf(px == 0 ? acc0[c + j] : acc1[c + j], k.scale[c + j], k.shift[c + j]);
        if (d.relu) y[j] = fmaxf(y[j], 0.f);
      }
      uint4 o;
      o.x = pack_act2(y[0], y[1], d.dtype);
      o.y = pack_act2(y[2], y[3], d.dtype);
      o.z = pack_act2(y[4], y[5], d.dtype);
      o.w = pack_act2(y[6], y[7], d.dtype);
      *reinterpret_cast<uint4*>(out + (long long)px * d.out_ld + c) = o;
    }
  }
}

// ------------------------------------------------------------------------------
// The same layer on the tensor cores: a [128 pixels] x [K = 32] x [32 channels] GEMM per tile whose A operand is
// built by the CTA itself straight from the u8 plane (no patch matrix in HBM).
//   K layout: the fp32 folded gray weight of every tap is split EXACTLY into three 16-bit terms
//   (w = hi + mid + lo: 3 x 8 significand bits for bf16 = all 24 of fp32; fp16: 33 bits),
//   k = term * 9 + tap (27 of 32 columns used), and the A row repeats the nine gray levels (integers <= 255: exact in
//   bf16 / fp16) three times.  Products are exact and accumulate in fp32 (TMEM), so the result equals the fp32 sum of
//   the FFMA kernel above up to the order of additions.
//   Roles (288 threads): warps 0-3 build A rows (one thread = one pixel: 9 byte loads -> 64-byte row, written in the
//   K-major SWIZZLE_64B layout the UMMA descriptor reads: 16-byte chunk c of row r at r*64 + ((c ^ (r>>1 & 3)) << 4));
//   warp 4 issues two tcgen05.mma (K = 16 each) per tile into one of two 32-column accumulators; warps 5-8 read TMEM,
//   apply shift / ReLU (the BN scale is folded into B), stage their 32 rows (2 KB, contiguous in the output tensor) and
//   hand them to ONE TMA store (cp.async.bulk.tensor shared -> global, 32 rows x 64 B).  Several CTAs per SM (44 KB smem, 64 TMEM columns each) overlap each other's phases; the kernel is
//   bound by the output write (64 B per pixel).
// ------------------------------------------------------------------------------
constexpr int kSuStages = 4;
constexpr int kSuThreads = 288;
constexpr int kSuATile = 128 * 64;
constexpr int kSuSmem = kSuStages * kSuATile + 2048 + 8 * 2048 + 16 * 8 + 1024;

using ptx::fence_proxy_async_smem;

// two fp32 -> packed 16-bit pair (lo, hi) in one F2FP; RELU / the fp16 saturation ride on the conversion
template <bool FP16, bool RELU>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  if (FP16) {
    if (RELU) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  }
  return r;
}
template <bool FP16>
__device__ __forceinline__ uint32_t to16(float v) {
  if (FP16) return (uint32_t)__half_as_ushort(__float2half_rn(v));
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
}
template <bool FP16>
__device__ __forceinline__ float from16(uint32_t u) {
  if (FP16) return __half2float(__ushort_as_half((unsigned short)u));
  return __uint_as_float(u << 16);
}

template <bool FP16, bool RELU>
__global__ void __launch_bounds__(kSuThreads) stem_gray3x3_umma_kernel(const ifcb_stem_desc d, const __grid_constant__ StemConst<32> k,
                                                                       const __grid_constant__ CUtensorMap tmap_out,
                                                                       int P, int Q, int total, int n_tiles, int tiles_per_cta,
                                                                       unsigned long long magic_pq, unsigned long long magic_q,
                                                                       float mul, float inv_mul) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_base = smem;                                   // kSuStages x 8 KB, SWIZZLE_64B rows
  uint8_t* b_tile = smem + kSuStages * kSuATile;            // 32 rows x 64 B
  uint8_t* s_stage = b_tile + 2048;                         // 4 epilogue warps x 2 buffers x 2 KB, SWIZZLE_64B rows (512-byte aligned)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_stage + 8 * 2048);   // [kSuStages] 128 producer arrivals
  uint64_t* a_empty = a_full + kSuStages;                         // [kSuStages] MMA commit
  uint64_t* acc_full = a_empty + kSuStages;                       // [2]
  uint64_t* acc_empty = acc_full + 2;                             // [2] 4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kSuStages; ++s) { ptx::mbar_init(a_full + s, 128); ptx::mbar_init(a_empty + s, 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(acc_full + s, 1); ptx::mbar_init(acc_empty + s, 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  if (tid < 32) {
    // B row `co`: [hi(9) | mid(9) | lo(9) | 0 x 5].  The BN scale is folded into the row, and the whole operand is scaled
    // by a power of two `mul` (undone exactly in the epilogue) that puts the largest entry near 2^14: the mid / lo terms
    // of small weights would otherwise fall into fp16's subnormal range
    const int co = tid;
    uint32_t e[32];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float w = k.w[t * 32 + co] * k.scale[co] * mul;
      const uint32_t h = to16<FP16>(w);
      const float r1 = w - from16<FP16>(h);
      const uint32_t m = to16<FP16>(r1);
      const float r2 = r1 - from16<FP16>(m);
      e[t] = h;
      e[9 + t] = m;
      e[18 + t] = to16<FP16>(r2);
    }
#pragma unroll
    for (int j = 27; j < 32; ++j) e[j] = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 v;
      v.x = e[8 * c + 0] | (e[8 * c + 1] << 16);
      v.y = e[8 * c + 2] | (e[8 * c + 3] << 16);
      v.z = e[8 * c + 4] | (e[8 * c + 5] << 16);
      v.w = e[8 * c + 6] | (e[8 * c + 7] << 16);
      ptx::st_shared_v4(ptx::smem_u32(b_tile) + co * 64 + ((c ^ ((co >> 1) & 3)) << 4), v);
    }
    fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // a CTA owns a contiguous range of tiles: pixel coordinates advance incrementally, and a warp's 32 output rows of
  // every tile are one contiguous 2 KB piece of the output tensor
  const int tile_begin = blockIdx.x * tiles_per_cta;
  const int tile_end = min(n_tiles, tile_begin + tiles_per_cta);

  if (warp < 4) {
    // ===================== A-row producers: one thread = one pixel of the tile =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t my_row = (uint32_t)tid * 64u;
    const uint32_t sw = ((uint32_t)tid >> 1) & 3u;
    const int step_w = d.stride, step_h = d.stride * d.W;
    // this thread's pixel of the first tile; the NEXT tile's is 128 pixels further (Q >= 128 is not assumed)
    int m = tile_begin * 128 + tid;
    int img = (int)fast_div((uint32_t)min(m, total - 1), magic_pq);
    int rem = min(m, total - 1) - img * (P * Q);
    int op = (int)fast_div((uint32_t)rem, magic_q);
    int oq = rem - op * Q;
    const int adv_p = 128 / Q, adv_q = 128 - adv_p * Q;             // 128 pixels = adv_p rows + adv_q columns
    auto load9 = [&](bool valid, uint32_t (&g)[9]) {
#pragma unroll
      for (int t = 0; t < 9; ++t) g[t] = 0;
      if (valid) {
        const uint8_t* in = reinterpret_cast<const uint8_t*>(d.d_in) + (long long)img * d.H * d.W + op * step_h + oq * step_w;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) g[r * 3 + c] = __ldg(in + r * d.W + c);
      }
    };
    auto advance = [&]() {
      m += 128;
      oq += adv_q;
      op += adv_p;
      if (oq >= Q) { oq -= Q; ++op; }
      while (op >= P) { op -= P; ++img; }
    };
    // the nine byte loads of the NEXT tile are in flight while this tile's row is converted and written
    uint32_t gnext[9];
    load9(tile_begin < tile_end && m < total, gnext);
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      float f[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) f[t] = (float)gnext[t];
      advance();
      load9(tile + 1 < tile_end && m < total, gnext);
      // k = term * 9 + tap, two k per word: ten distinct (tap, tap) pairs make the 27 used columns (one F2FP each, exact)
      const uint32_t p01 = pack2<FP16, false>(f[0], f[1]), p23 = pack2<FP16, false>(f[2], f[3]), p45 = pack2<FP16, false>(f[4], f[5]),
                     p67 = pack2<FP16, false>(f[6], f[7]), p80 = pack2<FP16, false>(f[8], f[0]), p12 = pack2<FP16, false>(f[1], f[2]),
                     p34 = pack2<FP16, false>(f[3], f[4]), p56 = pack2<FP16, false>(f[5], f[6]), p78 = pack2<FP16, false>(f[7], f[8]),
                     p8z = pack2<FP16, false>(f[8], 0.f);
      ptx::mbar_wait(a_empty + stage, phase ^ 1);
      const uint32_t row = ptx::smem_u32(a_base + (size_t)stage * kSuATile) + my_row;
      ptx::st_shared_v4(row + ((0u ^ sw) << 4), make_uint4(p01, p23, p45, p67));
      ptx::st_shared_v4(row + ((1u ^ sw) << 4), make_uint4(p80, p12, p34, p56));
      ptx::st_shared_v4(row + ((2u ^ sw) << 4), make_uint4(p78, p01, p23, p45));
      ptx::st_shared_v4(row + ((3u ^ sw) << 4), make_uint4(p67, p8z, 0u, 0u));
      fence_proxy_async_smem();
      ptx::mbar_arrive(a_full + stage);
      if (++stage == kSuStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ptx::umma_idesc_f16(128, 32, FP16 ? 1 : 0);
    const uint32_t desc_hi = ptx::umma_desc_hi(64);
    const uint32_t b_lo = ptx::umma_desc_lo(ptx::smem_u32(b_tile));
    const bool leader = ptx::elect_one();
    int stage = 0, local = 0;
    uint32_t phase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++local) {
      const int acc = local & 1;
      ptx::mbar_wait(acc_empty + acc, (uint32_t)(((local >> 1) & 1) ^ 1));
      ptx::mbar_wait(a_full + stage, phase);
      ptx::tc_fence_after();
      const uint32_t a_lo = ptx::umma_desc_lo(ptx::smem_u32(a_base + (size_t)stage * kSuATile));
      const uint32_t dt = tmem_base + (uint32_t)(acc * 32);
      if (leader) {
        ptx::umma_f16_lohi(dt, a_lo, b_lo, desc_hi, idesc, 0u);
        ptx::umma_f16_lohi(dt, a_lo + 2u, b_lo + 2u, desc_hi, idesc, 1u);
        ptx::umma_commit(a_empty + stage);
        ptx::umma_commit(acc_full + acc);
      }
      __syncwarp();
      if (++stage == kSuStages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== epilogue: TMEM -> shift / ReLU -> 16-bit -> staging -> one TMA store (32 rows x 64 B) per warp =====
    const int quad = warp & 3;                                  // the TMEM lane quadrant this warp may read
    const uint32_t stg0 = ptx::smem_u32(s_stage) + (uint32_t)(warp - 5) * 4096u;
    const uint32_t sw = ((uint32_t)lane >> 1) & 3u;             // same swizzle as the tensor map: conflict-free 16-byte stores
    if (lane == 0) ptx::prefetch_tensormap(&tmap_out);
    int local = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++local) {
      const int acc = local & 1;
      ptx::mbar_wait(acc_full + acc, (uint32_t)((local >> 1) & 1));
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 32);
      uint32_t v0[16], v1[16];
      ptx::tmem_ld_32x32b_x16(taddr, v0);
      ptx::tmem_ld_32x32b_x16(taddr + 16u, v1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      // the staging buffer written two tiles ago must have been read by its store
      if (lane == 0) ptx::bulk_wait_group_read<1>();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty + acc);
      const uint32_t stg = stg0 + (uint32_t)(local & 1) * 2048u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c0 = 8 * c + 2 * j;                     // compile-time: the shifts are constant-bank operands
          const float y0 = fmaf(__uint_as_float(c0 < 16 ? v0[c0 & 15] : v1[c0 & 15]), inv_mul, k.shift[c0]);
          const float y1 = fmaf(__uint_as_float(c0 < 16 ? v0[(c0 + 1) & 15] : v1[(c0 + 1) & 15]), inv_mul, k.shift[c0 + 1]);
          w[j] = pack2<FP16, RELU>(y0, y1);
        }
        ptx::st_shared_v4(stg + (uint32_t)lane * 64u + (((uint32_t)c ^ sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_2d(&tmap_out, stg, 0, tile * 128 + quad * 32);       // rows >= total are clipped
        ptx::bulk_commit_group();
      }
    }
    if (lane == 0) ptx::bulk_wait_group<0>();
    __syncwarp();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------
// 3x3 pooling over NHWC 16-bit (every pool of Inception-v3 / ResNet): one thread = one output
// pixel x 8 channels (16 bytes); index math by magic division, all nine 16-byte loads in
// flight before the reduction; max is taken on packed 16-bit pairs (exact), avg in fp32.
// ------------------------------------------------------------------------------
template <bool FP16>
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
  if (FP16) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

template <bool AVG, bool FP16>
__global__ void __launch_bounds__(256) pool3_kernel(const ifcb_pool_desc d, int P, int Q, int total, unsigned long long magic_c8,
                                                    unsigned long long magic_pq, unsigned long long magic_q) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int c8n = d.C >> 3;
  const int pix = (int)fast_div((uint32_t)idx, magic_c8);
  const int c8 = idx - pix * c8n;
  const int PQ = P * Q;
  const int img = (int)fast_div((uint32_t)pix, magic_pq);
  const int rem = pix - img * PQ;
  const int op = (int)fast_div((uint32_t)rem, magic_q), oq = rem - op * Q;
  const int h0 = op * d.stride - d.pad, w0 = oq * d.stride - d.pad;
  const int Wpi = d.W + 2 * d.in_pad_w;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(d.d_in) +
      ((long long)img * (d.H + 2 * d.in_pad_h) * Wpi + (long long)(d.in_pad_h + h0) * Wpi + d.in_pad_w + w0) * d.in_ld + c8 * 8;
  const uint32_t ninf = FP16 ? 0xFC00FC00u : 0xFF80FF80u;        // packed -inf pair
  const uint32_t fill = AVG ? 0u : ninf;
  uint4 v[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s2 = 0; s2 < 3; ++s2) {
      const bool ok = (unsigned)(h0 + r) < (unsigned)d.H && (unsigned)(w0 + s2) < (unsigned)d.W;
      v[r * 3 + s2] = ok ? __ldg(reinterpret_cast<const uint4*>(in + ((long long)r * Wpi + s2) * d.in_ld))
                         : make_uint4(fill, fill, fill, fill);
    }
  uint4 o;
  if (AVG) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const uint32_t u[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_act2(u[j], FP16 ? 1 : 0);
        a[2 * j] += f.x;
        a[2 * j + 1] += f.y;
      }
    }
    const float inv = 1.0f / 9.0f;                      // count_include_pad=True: always k*k
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(d.d_scale + c8 * 8)), s1 = __ldg(reinterpret_cast<const float4*>(d.d_scale + c8 * 8 + 4));
    const float4 h0v = __ldg(reinterpret_cast<const float4*>(d.d_shift + c8 * 8)), h1v = __ldg(reinterpret_cast<const float4*>(d.d_shift + c8 * 8 + 4));
    const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float shv[8] = {h0v.x, h0v.y, h0v.z, h0v.w, h1v.x, h1v.y, h1v.z, h1v.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float y = fmaf(a[j] * inv, scv[j], shv[j]);
      a[j] = d.relu ? fmaxf(y, 0.f) : y;
    }
    o.x = pack_act2(a[0], a[1], FP16 ? 1 : 0);
    o.y = pack_act2(a[2], a[3], FP16 ? 1 : 0);
    o.z = pack_act2(a[4], a[5], FP16 ? 1 : 0);
    o.w = pack_act2(a[6], a[7], FP16 ? 1 : 0);
  } else {
    o = v[0];
#pragma unroll
    for (int t = 1; t < 9; ++t) {
      o.x = max2<FP16>(o.x, v[t].x);
      o.y = max2<FP16>(o.y, v[t].y);
      o.z = max2<FP16>(o.z, v[t].z);
      o.w = max2<FP16>(o.w, v[t].w);
    }
  }
  const long long orow = ((long long)img * (P + 2 * d.out_pad_h) + op + d.out_pad_h) * (Q + 2 * d.out_pad_w) + oq + d.out_pad_w;
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.d_out) + orow * d.out_ld + c8 * 8) = o;
}

// ------------------------------------------------------------------------------
// Pooling over NHWC bf16: one thread = one output pixel x 8 channels (16 bytes).
// ------------------------------------------------------------------------------
template <bool AVG>
__global__ void __launch_bounds__(256) pool_kernel(const ifcb_pool_desc d, int P, int Q, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8n = d.C >> 3;
  const int c8 = (int)(idx % c8n);
  const long long pix = idx / c8n;
  const int PQ = P * Q;
  const int img = (int)(pix / PQ);
  const int rem = (int)(pix - (long long)img * PQ);
  const int op = rem / Q, oq = rem - op * Q;
  const int h0 = op * d.stride - d.pad, w0 = oq * d.stride - d.pad;
  const int Wpi = d.W + 2 * d.in_pad_w;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(d.d_in) +
      ((long long)img * (d.H + 2 * d.in_pad_h) * Wpi + (long long)d.in_pad_h * Wpi + d.in_pad_w) * d.in_ld + c8 * 8;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = AVG ? 0.f : -INFINITY;
  for (int r = 0; r < d.k; ++r) {
    const int hh = h0 + r;
    if (hh < 0 || hh >= d.H) continue;
    for (int s = 0; s < d.k; ++s) {
      const int ww = w0 + s;
      if (ww < 0 || ww >= d.W) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((long long)hh * Wpi + ww) * d.in_ld));
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_act2(u[j], d.dtype);
        if (AVG) {
          a[2 * j] += f.x;
          a[2 * j + 1] += f.y;
        } else {
          a[2 * j] = fmaxf(a[2 * j], f.x);
          a[2 * j + 1] = fmaxf(a[2 * j + 1], f.y);
        }
      }
    }
  }
  if (AVG) {
    const float inv = 1.0f / (float)(d.k * d.k);          // count_include_pad=True: always k*k
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
      float y = fmaf(a[j] * inv, d.d_scale[c], d.d_shift[c]);
      a[j] = d.relu ? fmaxf(y, 0.f) : y;
    }
  }
  uint4 o;
  o.x = pack_act2(a[0], a[1], d.dtype);
  o.y = pack_act2(a[2], a[3], d.dtype);
  o.z = pack_act2(a[4], a[5], d.dtype);
  o.w = pack_act2(a[6], a[7], d.dtype);
  const long long orow = ((long long)img * (P + 2 * d.out_pad_h) + op + d.out_pad_h) * (Q + 2 * d.out_pad_w) + oq + d.out_pad_w;
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.d_out) + orow * d.out_ld + c8 * 8) = o;
}

// ------------------------------------------------------------------------------
// Head: one CTA per image.  (1) spatial mean per channel (fp32) into shared
// memory, (2) logits: one warp per class, float4 weight loads + shuffle
// reduction, (3) numerically stable softmax + first-index argmax by warp 0.
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_kernel(const ifcb_head_desc d) {
  extern __shared__ float sm[];
  float* pooled = sm;                    // [C]
  float* logits = sm + d.C;              // [n_classes]
  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(d.d_in) + (long long)img * d.HW * d.in_ld;
  const float inv = 1.0f / (float)d.HW;
  for (int c8 = tid; c8 < (d.C >> 3); c8 += blockDim.x) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int px = 0; px < d.HW; ++px) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (long long)px * d.in_ld + c8 * 8));
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_act2(u[j], d.dtype);
        a[2 * j] += f.x;
        a[2 * j + 1] += f.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) pooled[c8 * 8 + j] = a[j] * inv;
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int k = warp; k < d.n_classes; k += nw) {
    const float4* wrow = reinterpret_cast<const float4*>(d.d_weight + (long long)k * d.C);
    float s = 0.f;
    for (int c4 = lane; c4 < (d.C >> 2); c4 += 32) {
      const float4 w = __ldg(wrow + c4);
      const float4 x = *reinterpret_cast<const float4*>(pooled + c4 * 4);
      s = fmaf(w.x, x.x, s);
      s = fmaf(w.y, x.y, s);
      s = fmaf(w.z, x.z, s);
      s = fmaf(w.w, x.w, s);
    }
    s = warp_sum(s);
    if (lane == 0) logits[k] = s + d.d_bias[k];
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int k = lane; k < d.n_classes; k += 32) mx = fmaxf(mx, logits[k]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < d.n_classes; k += 32) sum += expf(logits[k] - mx);
    sum = warp_sum(sum);
    float best = -1.f;
    int best_k = 0x7fffffff;
    for (int k = lane; k < d.n_classes; k += 32) {
      const float l = logits[k];
      const float pr = expf(l - mx) / sum;
      d.d_scores[(long long)img * d.n_classes + k] = pr;
      if (d.d_logits) d.d_logits[(long long)img * d.n_classes + k] = l;
      if (pr > best) { best = pr; best_k = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
      if (ob > best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
    }
    if (lane == 0) {
      d.d_top1[img] = best_k;
      d.d_top1_score[img] = best;
    }
  }
}

// ------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 average pool (+ BN affine + ReLU): the Inception branch_pool path.
// One thread = FOUR horizontally adjacent output pixels x 8 channels: the 6 x 3 input window is
// loaded and converted once, summed by column, and each output adds three column sums --
// half the loads, conversions and adds of one-output-per-thread.  fp32 accumulation;
// count_include_pad=True (torch default, inception.py:271): always divide by 9.
// ------------------------------------------------------------------------------
template <bool FP16>
__global__ void __launch_bounds__(256) avgpool3_s1_kernel(const ifcb_pool_desc d, int total, int q4n, unsigned long long magic_c8,
                                                          unsigned long long magic_q4, unsigned long long magic_h) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int c8n = d.C >> 3;
  const int t0 = (int)fast_div((uint32_t)idx, magic_c8);
  const int c8 = idx - t0 * c8n;
  const int t1 = (int)fast_div((uint32_t)t0, magic_q4);
  const int q0 = (t0 - t1 * q4n) * 4;                 // first output column of this thread
  const int img = (int)fast_div((uint32_t)t1, magic_h);
  const int op = t1 - img * d.H;
  const int Wpi = d.W + 2 * d.in_pad_w;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(d.d_in) +
      ((long long)img * (d.H + 2 * d.in_pad_h) * Wpi + (long long)(d.in_pad_h + op - 1) * Wpi + d.in_pad_w + q0 - 1) * d.in_ld + c8 * 8;
  float col[6][8];
#pragma unroll
  for (int cidx = 0; cidx < 6; ++cidx) {
#pragma unroll
    for (int j = 0; j < 8; ++j) col[cidx][j] = 0.f;
    const int ww = q0 - 1 + cidx;
    const bool wok = (unsigned)ww < (unsigned)d.W;
    uint4 v[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const bool ok = wok && (unsigned)(op - 1 + r) < (unsigned)d.H;
      v[r] = ok ? __ldg(reinterpret_cast<const uint4*>(in + ((long long)r * Wpi + cidx) * d.in_ld)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const uint32_t u[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_act2(u[j], FP16 ? 1 : 0);
        col[cidx][2 * j] += f.x;
        col[cidx][2 * j + 1] += f.y;
      }
    }
  }
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(d.d_scale + c8 * 8)), s1 = __ldg(reinterpret_cast<const float4*>(d.d_scale + c8 * 8 + 4));
  const float4 h0 = __ldg(reinterpret_cast<const float4*>(d.d_shift + c8 * 8)), h1 = __ldg(reinterpret_cast<const float4*>(d.d_shift + c8 * 8 + 4));
  const float inv = 1.0f / 9.0f;
  const float scv[8] = {s0.x * inv, s0.y * inv, s0.z * inv, s0.w * inv, s1.x * inv, s1.y * inv, s1.z * inv, s1.w * inv};
  const float shv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.d_out) +
      (((long long)img * (d.H + 2 * d.out_pad_h) + op + d.out_pad_h) * (d.W + 2 * d.out_pad_w) + d.out_pad_w + q0) * d.out_ld + c8 * 8;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    if (q0 + o < d.W) {
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y[j] = fmaf(col[o][j] + col[o + 1][j] + col[o + 2][j], scv[j], shv[j]);
        if (d.relu) y[j] = fmaxf(y[j], 0.f);
      }
      uint4 w;
      w.x = pack_act2(y[0], y[1], FP16 ? 1 : 0);
      w.y = pack_act2(y[2], y[3], FP16 ? 1 : 0);
      w.z = pack_act2(y[4], y[5], FP16 ? 1 : 0);
      w.w = pack_act2(y[6], y[7], FP16 ? 1 : 0);
      *reinterpret_cast<uint4*>(out + (long long)o * d.out_ld) = w;
    }
  }
}

}  // namespace

int launch_stem(const StemLayer& L, int batch, cudaStream_t stream) {
  const ifcb_stem_desc& d = L.d;
  const long long total = (long long)batch * L.P * L.Q;
  if (total == 0) return 0;
  const int taps = d.kh * d.kw;
  const unsigned grid = (unsigned)((total + 127) / 128);
  if (d.Cout != 32 && d.Cout != 64 && d.Cout != 96) {
    set_error("stem: Cout=%d unsupported (32, 64 or 96)", d.Cout);
    return -1;
  }
  if (d.in_kind == IFCB_STEM_IN_U8_GRAY && d.kh == 3 && d.kw == 3 && d.pad == 0 && (d.stride == 1 || d.stride == 2) && d.Cout == 32 &&
      total < (1ll << 31) && !L.h_const.empty()) {
    // tensor-core version (A/B switch IFCB_STEM_UMMA=0: the FFMA kernel)
    static const bool umma = !(getenv("IFCB_STEM_UMMA") && atoi(getenv("IFCB_STEM_UMMA")) == 0);
    const long long opix_max = (long long)batch * (L.P + 2 * d.out_pad_h) * (L.Q + 2 * d.out_pad_w);
    if (umma && opix_max < (1ll << 31) && d.out_ld == 32 && d.out_pad_h == 0 && d.out_pad_w == 0 && L.P * L.Q >= 128 &&
        (reinterpret_cast<uintptr_t>(d.d_out) & 15) == 0) {
      StemConst<32> k;
      memcpy(&k, L.h_const.data(), sizeof(k));
      const int n_tiles = (int)((total + 127) / 128);
      // one wave of resident CTAs (a CTA's tile loop is latency-bound; a partial second wave would run at a fraction of the rate)
      void (*kern)(const ifcb_stem_desc, const StemConst<32>, const CUtensorMap, int, int, int, int, int, unsigned long long, unsigned long long, float, float) =
          d.dtype ? (d.relu ? stem_gray3x3_umma_kernel<true, true> : stem_gray3x3_umma_kernel<true, false>)
                  : (d.relu ? stem_gray3x3_umma_kernel<false, true> : stem_gray3x3_umma_kernel<false, false>);
      static int per_sm[64][4] = {};
      int dev = 0;
      IFCB_CUDA_CHECK(cudaGetDevice(&dev));
      const int variant = (d.dtype ? 2 : 0) + (d.relu ? 1 : 0);
      int occ = (dev >= 0 && dev < 64) ? per_sm[dev][variant] : 0;
      if (occ == 0) {
        cudaFuncAttributes fa;
        IFCB_CUDA_CHECK(cudaFuncGetAttributes(&fa, kern));
        IFCB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSuSmem));
        IFCB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        const int warps = kSuThreads / 32;
        const int regs_per_warp = ((fa.numRegs * 32 + 255) / 256) * 256;
        const int by_regs = 65536 / (regs_per_warp * warps);
        const int by_smem = (227 * 1024) / (kSuSmem + 1024);
        occ = by_regs < by_smem ? by_regs : by_smem;
        if (occ > 2048 / kSuThreads) occ = 2048 / kSuThreads;
        if (occ > 8) occ = 8;                       // 64 TMEM columns per CTA
        if (occ < 1) occ = 1;
        if (dev >= 0 && dev < 64) per_sm[dev][variant] = occ;
      }
      int ctas = n_tiles < occ * sm_count() ? n_tiles : occ * sm_count();
      const int tiles_per_cta = (n_tiles + ctas - 1) / ctas;
      ctas = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
      // power of two that puts the largest folded weight (w * scale) in [2^14, 2^15)
      float wmax = 0.f;
      for (int t = 0; t < 9; ++t)
        for (int c = 0; c < 32; ++c) {
          const float v = fabsf(k.w[t * 32 + c] * k.scale[c]);
          if (v > wmax && v < 3.0e38f) wmax = v;
        }
      int ex = 0;
      if (wmax > 0.f) frexpf(wmax, &ex);
      int up = 15 - ex;
      if (up > 100) up = 100;
      if (up < -100) up = -100;
      // the output as a [total pixels] x [32 channels] matrix, stored in 32-row boxes (SWIZZLE_64B: the staging layout)
      CUtensorMap tmap_out;
      {
        int rc = resolve_driver();
        if (rc) return rc;
        cuuint64_t gdim[2] = {32, (cuuint64_t)total};
        cuuint64_t gstr[1] = {64};
        cuuint32_t box[2] = {32, 32};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encode_tiled(&tmap_out, d.dtype ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d.d_out, gdim, gstr,
                                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
          set_error("stem: cuTensorMapEncodeTiled (output) failed (%d)", (int)r);
          return -1;
        }
      }
      kern<<<ctas, kSuThreads, kSuSmem, stream>>>(d, k, tmap_out, L.P, L.Q, (int)total, n_tiles, tiles_per_cta, div_magic(L.P * L.Q), div_magic(L.Q), ldexpf(1.f, up),
                                                  ldexpf(1.f, -up));
      IFCB_CUDA_CHECK(cudaGetLastError());
      return 0;
    }
    const int q2n = (L.Q + 1) / 2;
    const long long t2 = (long long)batch * L.P * q2n;
    const unsigned blocks = (unsigned)((t2 + 127) / 128);
    StemConst<32> k;
    memcpy(&k, L.h_const.data(), sizeof(k));
    stem_gray3x3_kernel<32><<<blocks, 128, 0, stream>>>(d, k, L.P, L.Q, q2n, (int)t2, div_magic(L.P * q2n), div_magic(q2n));
  } else if (d.in_kind == IFCB_STEM_IN_U8_GRAY) {
    const bool haspad = d.pad > 0;
    const int smem = taps * d.Cout * (haspad ? 2 : 1) * (int)sizeof(float);
#define IFCB_GRAY_LAUNCH(CO, HP)                                                                          \
  do {                                                                                                    \
    if (smem > 48 * 1024)                                                                                 \
      IFCB_CUDA_CHECK(cudaFuncSetAttribute(stem_gray_kernel<CO, HP>,                                      \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem));           \
    stem_gray_kernel<CO, HP><<<grid, 128, smem, stream>>>(d, L.P, L.Q, total);                            \
  } while (0)
    if (d.Cout == 32) {
      if (haspad) IFCB_GRAY_LAUNCH(32, true); else IFCB_GRAY_LAUNCH(32, false);
    } else if (d.Cout == 64) {
      if (haspad) IFCB_GRAY_LAUNCH(64, true); else IFCB_GRAY_LAUNCH(64, false);
    } else {                      // densenet161
      if (haspad) IFCB_GRAY_LAUNCH(96, true); else IFCB_GRAY_LAUNCH(96, false);
    }
#undef IFCB_GRAY_LAUNCH
  } else {
    const int smem = (taps * 3 * d.Cout + 768) * (int)sizeof(float);
#define IFCB_STEM_LAUNCH(CO)                                                                              \
  do {                                                                                                    \
    if (smem > 48 * 1024)                                                                                 \
      IFCB_CUDA_CHECK(cudaFuncSetAttribute(stem_kernel<CO, false>,                                        \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem));           \
    stem_kernel<CO, false><<<grid, 128, smem, stream>>>(d, nullptr, L.P, L.Q, total);                     \
  } while (0)
    if (d.Cout == 32) IFCB_STEM_LAUNCH(32); else if (d.Cout == 64) IFCB_STEM_LAUNCH(64); else IFCB_STEM_LAUNCH(96);
#undef IFCB_STEM_LAUNCH
  }
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_pool(const PoolLayer& L, int batch, cudaStream_t stream) {
  const ifcb_pool_desc& d = L.d;
  const long long total = (long long)batch * L.P * L.Q * (d.C >> 3);
  if (total == 0) return 0;
  const unsigned grid = (unsigned)((total + 255) / 256);
  const bool avg = d.kind != IFCB_POOL_MAX;
  if (avg && d.k == 3 && d.stride == 1 && d.pad == 1 && total < (1ll << 31)) {
    const int q4n = (d.W + 3) / 4;
    const long long t4 = (long long)batch * d.H * q4n * (d.C >> 3);
    const unsigned g4 = (unsigned)((t4 + 255) / 256);
    if (d.dtype) avgpool3_s1_kernel<true><<<g4, 256, 0, stream>>>(d, (int)t4, q4n, div_magic(d.C >> 3), div_magic(q4n), div_magic(d.H));
    else avgpool3_s1_kernel<false><<<g4, 256, 0, stream>>>(d, (int)t4, q4n, div_magic(d.C >> 3), div_magic(q4n), div_magic(d.H));
  } else if (d.k == 3 && total < (1ll << 31)) {
    const unsigned long long mc = div_magic(d.C >> 3), mpq = div_magic(L.P * L.Q), mq = div_magic(L.Q);
    const int t = (int)total;
    if (avg) {
      if (d.dtype) pool3_kernel<true, true><<<grid, 256, 0, stream>>>(d, L.P, L.Q, t, mc, mpq, mq);
      else pool3_kernel<true, false><<<grid, 256, 0, stream>>>(d, L.P, L.Q, t, mc, mpq, mq);
    } else {
      if (d.dtype) pool3_kernel<false, true><<<grid, 256, 0, stream>>>(d, L.P, L.Q, t, mc, mpq, mq);
      else pool3_kernel<false, false><<<grid, 256, 0, stream>>>(d, L.P, L.Q, t, mc, mpq, mq);
    }
  } else if (!avg) {
    pool_kernel<false><<<grid, 256, 0, stream>>>(d, L.P, L.Q, total);
  } else {
    pool_kernel<true><<<grid, 256, 0, stream>>>(d, L.P, L.Q, total);
  }
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_head(const HeadLayer& L, int batch, int out_row, cudaStream_t stream) {
  ifcb_head_desc d = L.d;
  if (batch == 0) return 0;
  if (out_row) {                  // rows of this batch land at out_row of the caller's per-bin output buffers
    d.d_scores = static_cast<float*>(d.d_scores) + (size_t)out_row * d.n_classes;
    if (d.d_logits) d.d_logits = static_cast<float*>(d.d_logits) + (size_t)out_row * d.n_classes;
    d.d_top1 = static_cast<int32_t*>(d.d_top1) + out_row;
    d.d_top1_score = static_cast<float*>(d.d_top1_score) + out_row;
  }
  const int smem = (d.C + d.n_classes) * (int)sizeof(float);
  head_kernel<<<batch, 256, smem, stream>>>(d);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace ifcb
