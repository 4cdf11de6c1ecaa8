// Stem convolution (Cin = 3, direct fp32), pooling (K3/K4) and the fused head
// (K6: global average pool -> Linear -> softmax -> top-1).  All HBM-bound
// integer/elementwise work: coalesced 16-byte accesses over NHWC bf16, warp
// shuffles for the reductions.  See include/ifcb_b200.h for the reference code
// each replaces.
#include "layers.cuh"
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
