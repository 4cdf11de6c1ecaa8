// K1 -- fused IFCB ROI preprocess (sm_100a).
//
// Replaces IfcbBinDataset.__getitem__ (reference neuston_data.py:456-464):
//   ToPILImage('L') -> convert('RGB') -> Resize((R,R)) -> ToTensor() -> [Normalize]
// One CTA per ROI.  The raw ROI bytes (a contiguous h*w block of the .roi file) are copied into
// shared memory with 16-byte vector loads and resampled with Pillow's two separable fixed-point
// passes (Resample.c, 8bpc, triangle filter, 22-bit coefficients generated in IEEE double exactly
// as precompute_coeffs does) entirely in shared memory:
//   * horizontal pass: a thread owns an output column, keeps its (<= 4) coefficients in registers
//     and walks down the input rows;
//   * vertical pass: a thread owns FOUR adjacent output pixels of a row -- one 32-bit shared load
//     per filter tap feeds four multiply-adds, the tap coefficient is shared by the whole row;
//   * the resampled gray bytes of up to 32 output rows are staged LINEARLY in shared memory and
//     leave the SM as 16-byte vector stores whatever the alignment of the output row (R = 299 is
//     odd): u8 plane as is, f32 / bf16 NCHW through the ToTensor/Normalize LUT per channel.
// A ROI whose rows all fit is resampled horizontally once ("resident"); larger ROIs are processed
// in bands of output rows; bounds too large to stage input rows at all (photos of several thousand
// pixels through `--type img` / TRAIN) read the horizontal pass straight from global memory.
//
// HBM traffic per ROI = h*w bytes in + one write of the output tensor.
#include "common.cuh"
#include "../../include/ifcb_b200.h"

namespace ifcb {
namespace {

constexpr int kThreads = 256;
constexpr int kPrecisionBits = 22;          // Resample.c: 32 - 8 - 2
constexpr int kOutBand = 32;                // output rows staged per write-out
constexpr int kSmemTarget = 56 * 1024;      // 4 CTAs / SM whenever the declared ROI bound allows it
constexpr int kSmemMax = 227 * 1024;

struct PreParams {
  const uint8_t* packed;
  long long packed_bytes;
  const long long* offsets;
  const int* hs;
  const int* ws;
  int n, R;
  int out_mode, pass_rule;
  int has_norm;
  float mean[3], stdv[3];
  void* out;
  int smem_bytes;      // dynamic shared memory of this launch
  int direct;          // 1: the horizontal pass reads the ROI from global memory (no input staging)
  int* status;         // optional device word: OR of IFCB_PRE_* flags of ROIs that could not be processed
};

// ksize of precompute_coeffs for the triangle filter (support 1.0).
__host__ __device__ inline int resample_ksize(int in_size, int out_size) {
  double scale = (double)in_size / (double)out_size;
  double fs = scale < 1.0 ? 1.0 : scale;
  return (int)ceil(fs) * 2 + 1;
}

// Coefficients of one axis: for output xx, taps [xmin, xmin+cnt) packed as tab[xx] = xmin | cnt << 20,
// int coeffs kk[xx*ksize + k].  Restates precompute_coeffs + normalize_coeffs_8bpc (no FMA contraction: _rn intrinsics).
__device__ void gen_coeffs(int in_size, int out_size, int ksize, int* __restrict__ kk, int* __restrict__ tab) {
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = fs;                 // 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, fs);
  for (int xx = threadIdx.x; xx < out_size; xx += blockDim.x) {
    const double center = __dadd_rn(0.0, __dmul_rn((double)xx + 0.5, scale));
    int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (xmax > in_size) xmax = in_size;
    const int cnt = xmax - xmin;
    double ww = 0.0;
    for (int x = 0; x < cnt; ++x) {
      double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
      if (a < 0.0) a = -a;
      const double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
      ww = __dadd_rn(ww, w);
    }
    int* k = kk + xx * ksize;
    for (int x = 0; x < cnt; ++x) {
      double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
      if (a < 0.0) a = -a;
      double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      k[x] = __double2int_rz(__dadd_rn(0.5, __dmul_rn(w, (double)(1 << kPrecisionBits))));
    }
    for (int x = cnt; x < ksize; ++x) k[x] = 0;
    tab[xx] = xmin | (cnt << 20);
  }
}

__device__ __forceinline__ int clip8(int acc) {
  int v = acc >> kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__device__ __forceinline__ void store_pixel(const PreParams& p, const float* lut, long long roi,
                                            int y, int x, int v) {
  const int R = p.R;
  if (p.out_mode == IFCB_OUT_U8_GRAY) {
    reinterpret_cast<uint8_t*>(p.out)[roi * R * R + (long long)y * R + x] = (uint8_t)v;
  } else if (p.out_mode == IFCB_OUT_F32_NCHW) {
    float* o = reinterpret_cast<float*>(p.out) + roi * 3 * R * R + (long long)y * R + x;
    o[0] = lut[v];
    o[(long long)R * R] = lut[256 + v];
    o[2ll * R * R] = lut[512 + v];
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + roi * 3 * R * R + (long long)y * R + x;
    o[0] = __float2bfloat16_rn(lut[v]);
    o[(long long)R * R] = __float2bfloat16_rn(lut[256 + v]);
    o[2ll * R * R] = __float2bfloat16_rn(lut[512 + v]);
  }
}

// 4 consecutive bytes at an arbitrary byte offset of a 4-byte aligned shared array
__device__ __forceinline__ uint32_t lds_u32_unaligned(const uint8_t* base, int off) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (off >> 2);
  return __funnelshift_r(w[0], w[1], (off & 3) * 8);
}

// Linear write-out of `len` staged gray bytes (stage[0 .. len)) to the output rows they cover.
// The destination is contiguous (rows of one plane follow each other), so the body is 16-byte
// vector stores regardless of R; head / tail elements up to the alignment boundary go one by one.
__device__ void write_out(const PreParams& p, const uint8_t* __restrict__ stage, const float* __restrict__ lut,
                          long long roi, int y0, int len) {
  const int R = p.R, tid = threadIdx.x;
  const long long pix0 = (long long)y0 * R;
  if (p.out_mode == IFCB_OUT_U8_GRAY) {
    uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + roi * R * R + pix0;
    int head = (int)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
    if (head > len) head = len;
    const int nv = (len - head) >> 4;
    if (tid < head) dst[tid] = stage[tid];
    for (int j = tid; j < nv; j += kThreads) {
      const int s = head + 16 * j;
      uint4 v;
      v.x = lds_u32_unaligned(stage, s);
      v.y = lds_u32_unaligned(stage, s + 4);
      v.z = lds_u32_unaligned(stage, s + 8);
      v.w = lds_u32_unaligned(stage, s + 12);
      *reinterpret_cast<uint4*>(dst + s) = v;
    }
    const int t0 = head + 16 * nv;
    if (tid < len - t0) dst[t0 + tid] = stage[t0 + tid];
    return;
  }
  for (int c = 0; c < 3; ++c) {
    const float* l = lut + 256 * c;
    if (p.out_mode == IFCB_OUT_F32_NCHW) {
      float* dst = reinterpret_cast<float*>(p.out) + (roi * 3 + c) * R * R + pix0;
      int head = (int)(((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15) >> 2);
      if (head > len) head = len;
      const int nv = (len - head) >> 2;
      if (tid < head) dst[tid] = l[stage[tid]];
      for (int j = tid; j < nv; j += kThreads) {
        const int s = head + 4 * j;
        const uint32_t b = lds_u32_unaligned(stage, s);
        float4 v;
        v.x = l[b & 255u];
        v.y = l[(b >> 8) & 255u];
        v.z = l[(b >> 16) & 255u];
        v.w = l[b >> 24];
        *reinterpret_cast<float4*>(dst + s) = v;
      }
      const int t0 = head + 4 * nv;
      if (tid < len - t0) dst[t0 + tid] = l[stage[t0 + tid]];
    } else {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + (roi * 3 + c) * R * R + pix0;
      int head = (int)(((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15) >> 1);
      if (head > len) head = len;
      const int nv = (len - head) >> 3;
      if (tid < head) dst[tid] = __float2bfloat16_rn(l[stage[tid]]);
      for (int j = tid; j < nv; j += kThreads) {
        const int s = head + 8 * j;
        const uint32_t b0 = lds_u32_unaligned(stage, s), b1 = lds_u32_unaligned(stage, s + 4);
        uint4 v;
        v.x = pack_bf16x2(l[b0 & 255u], l[(b0 >> 8) & 255u]);
        v.y = pack_bf16x2(l[(b0 >> 16) & 255u], l[b0 >> 24]);
        v.z = pack_bf16x2(l[b1 & 255u], l[(b1 >> 8) & 255u]);
        v.w = pack_bf16x2(l[(b1 >> 16) & 255u], l[b1 >> 24]);
        *reinterpret_cast<uint4*>(dst + s) = v;
      }
      const int t0 = head + 8 * nv;
      if (tid < len - t0) dst[t0 + tid] = __float2bfloat16_rn(l[stage[t0 + tid]]);
    }
  }
}

// A ROI that cannot be processed (table entry outside the packed buffer, or larger than the bound the launch was sized
// for): its output slot is ZEROED (never left holding a previous batch's pixels) and the status word records why.
__device__ void reject_roi(const PreParams& p, long long roi, int flag) {
  const long long n = (long long)p.R * p.R * (p.out_mode == IFCB_OUT_U8_GRAY ? 1 : 3);
  if (p.out_mode == IFCB_OUT_U8_GRAY) {
    uint8_t* o = reinterpret_cast<uint8_t*>(p.out) + roi * n;
    for (long long i = threadIdx.x; i < n; i += kThreads) o[i] = 0;
  } else if (p.out_mode == IFCB_OUT_F32_NCHW) {
    float* o = reinterpret_cast<float*>(p.out) + roi * n;
    for (long long i = threadIdx.x; i < n; i += kThreads) o[i] = 0.f;
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + roi * n;
    for (long long i = threadIdx.x; i < n; i += kThreads) o[i] = __float2bfloat16_rn(0.f);
  }
  if (threadIdx.x == 0 && p.status != nullptr) atomicOr(p.status, flag);
}

__global__ void __launch_bounds__(kThreads) preprocess_kernel(const PreParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int roi = blockIdx.x;
  const int R = p.R;
  const int h = p.hs[roi], w = p.ws[roi];
  const long long off = p.offsets[roi];
  const int tid = threadIdx.x;
  if (h <= 0 || w <= 0 || off < 0 || off + (long long)h * w > p.packed_bytes) {
    reject_roi(p, roi, IFCB_PRE_BAD_TABLE);
    return;
  }

  // ---- shared memory carve-up (tables sized by THIS ROI's filter supports) --------------------
  float* lut = reinterpret_cast<float*>(smem);             // [3][256]
  int* tab_x = reinterpret_cast<int*>(lut + 768);          // [R]  xmin | cnt << 20
  int* tab_y = tab_x + R;                                   // [R]
  const int ksx = resample_ksize(w, R), ksy = resample_ksize(h, R);
  int* kx = tab_y + R;                                      // [R*ksx]
  int* ky = kx + R * ksx;                                   // [R*ksy]
  int* misc = ky + R * ksy;                                 // [4] band bookkeeping
  uint8_t* stage = smem + ((reinterpret_cast<uint8_t*>(misc + 4) - smem + 15) & ~(size_t)15);    // [kOutBand*R] linear gray bytes
  uint8_t* bufs = stage + ((kOutBand * R + 4 + 15) & ~15);
  const int avail = p.smem_bytes - (int)(bufs - smem);
  const int inter_stride = (R + 3) & ~3;
  const bool direct = p.direct != 0;
  // rows of input a band may hold: staged input needs rows*w + 48 (alignment slack), the intermediate rows*inter_stride
  const int row_cost = (direct ? 0 : w) + inter_stride;
  int rows_cap = (avail - 64) / row_cost;
  if (rows_cap < ksy) {                                     // larger than the bound this launch was sized for
    reject_roi(p, roi, IFCB_PRE_TOO_LARGE);
    return;
  }
  const bool resident = rows_cap >= h;                      // every input row fits: one horizontal pass for the whole ROI
  if (rows_cap > h) rows_cap = h;

  // ToTensor + Normalize as a 256-entry LUT per channel (float32, same rounding sequence as torch: x/255, then (x-mean)/std)
  if (p.out_mode != IFCB_OUT_U8_GRAY) {
    for (int i = tid; i < 768; i += kThreads) {
      const int c = i >> 8, g = i & 255;
      float v = __fdiv_rn((float)g, 255.0f);
      if (p.has_norm) v = __fdiv_rn(__fsub_rn(v, p.mean[c]), p.stdv[c]);
      lut[i] = v;
    }
  }
  gen_coeffs(w, R, ksx, kx, tab_x);
  gen_coeffs(h, R, ksy, ky, tab_y);
  __syncthreads();

  const uint8_t* __restrict__ src = p.packed + off;

  // ---- Pillow >= 12 sliver rule: vertical pass first (w <= 10: rare, simple path) -------------
  if (p.pass_rule == IFCB_PASS_PILLOW12 && h > R && (long long)h > 100ll * w && w != R) {
    uint8_t* inter = bufs;                                  // [R][w]
    for (int idx = tid; idx < R * w; idx += kThreads) {
      const int y = idx / w, x = idx - y * w;
      const int* k = ky + y * ksy;
      const int y0 = tab_y[y] & 0xFFFFF, c = tab_y[y] >> 20;
      int acc = 1 << (kPrecisionBits - 1);
      for (int t = 0; t < c; ++t) acc += k[t] * (int)src[(long long)(y0 + t) * w + x];
      inter[idx] = (uint8_t)clip8(acc);
    }
    __syncthreads();
    for (int idx = tid; idx < R * R; idx += kThreads) {
      const int y = idx / R, x = idx - y * R;
      const int* k = kx + x * ksx;
      const int x0 = tab_x[x] & 0xFFFFF, c = tab_x[x] >> 20;
      int acc = 1 << (kPrecisionBits - 1);
      for (int t = 0; t < c; ++t) acc += k[t] * (int)inter[y * w + x0 + t];
      store_pixel(p, lut, roi, y, x, clip8(acc));
    }
    return;
  }

  // thread mapping of the vertical pass: tx = group of 4 output columns, ty = row lane
  const int x4n = (R + 3) >> 2;
  const int lanes_y = kThreads / x4n;                       // R <= 512  =>  x4n <= 128  =>  lanes_y >= 2
  const int tx = tid % x4n, ty = tid / x4n;

  int y0 = 0, have_r0 = 0, have_r1 = 0;                     // input rows [have_r0, have_r1) are resampled horizontally in `inter`
  uint8_t* inter = nullptr;
  while (y0 < R) {
    // thread 0 picks the band [y0, y1): at most kOutBand output rows whose input rows fit rows_cap
    if (tid == 0) {
      const int r0 = resident ? 0 : (tab_y[y0] & 0xFFFFF);
      int y1 = y0 + 1;
      const int ylim = min(R, y0 + kOutBand);
      while (y1 < ylim && (resident || (tab_y[y1] & 0xFFFFF) + (tab_y[y1] >> 20) - r0 <= rows_cap)) ++y1;
      misc[0] = y1;
      misc[1] = r0;
      misc[2] = resident ? h : (tab_y[y1 - 1] & 0xFFFFF) + (tab_y[y1 - 1] >> 20);
    }
    __syncthreads();
    const int y1 = misc[0], r0 = misc[1], r1 = misc[2];
    const int nrows = r1 - r0;

    if (!(resident && have_r1 > 0)) {
      // (1) contiguous vectorised copy of rows [r0, r1) into shared memory (skipped in direct mode)
      const long long gbeg = off + (long long)r0 * w;         // byte offset in packed
      const long long gend = off + (long long)r1 * w;
      const long long abeg = gbeg & ~15ll;
      const int head = (int)(gbeg - abeg);
      uint8_t* in_band = bufs;                                 // 16-byte aligned
      const int nvec = direct ? 0 : (int)((gend - abeg + 15) >> 4);
      inter = bufs + (((size_t)nvec * 16 + 15) & ~(size_t)15) + (direct ? 0 : 16);
      for (int i = tid; i < nvec; i += kThreads) {
        const long long g = abeg + 16ll * i;
        uint4 v;
        if (g + 16 <= p.packed_bytes) {
          v = __ldg(reinterpret_cast<const uint4*>(p.packed + g));
        } else {
          uint8_t tmp[16];
#pragma unroll
          for (int b = 0; b < 16; ++b) tmp[b] = (g + b < p.packed_bytes) ? p.packed[g + b] : 0;
          v = *reinterpret_cast<uint4*>(tmp);
        }
        reinterpret_cast<uint4*>(in_band)[i] = v;
      }
      __syncthreads();

      // (2) horizontal pass: a thread owns output column xo and walks down the rows; its coefficients sit in registers
      // (pointers into shared and global memory are kept apart so that the staged path compiles to LDS, not generic loads)
      const uint8_t* inb = in_band + head;
      const uint8_t* __restrict__ ing = src + (long long)r0 * w;
      for (int xo = tid; xo < R; xo += kThreads) {
        const int tb = tab_x[xo];
        const int x0 = tb & 0xFFFFF, c = tb >> 20;
        const int* k = kx + xo * ksx;
        uint8_t* dst = inter + xo;
        if (direct) {
          const uint8_t* row = ing + x0;
          for (int r = 0; r < nrows; ++r, row += w, dst += inter_stride) {
            int acc = 1 << (kPrecisionBits - 1);
            for (int t = 0; t < c; ++t) acc += k[t] * (int)__ldg(row + t);
            *dst = (uint8_t)clip8(acc);
          }
        } else if (c <= 4) {
          // taps past c carry a zero coefficient; their byte reads stay inside the staged band (16 bytes of slack behind it)
          const int k0 = k[0], k1 = c > 1 ? k[1] : 0, k2 = c > 2 ? k[2] : 0, k3 = c > 3 ? k[3] : 0;
          const uint8_t* row = inb + x0;
#pragma unroll 2
          for (int r = 0; r < nrows; ++r, row += w, dst += inter_stride) {
            int acc = 1 << (kPrecisionBits - 1);
            acc += k0 * (int)row[0];
            acc += k1 * (int)row[1];
            acc += k2 * (int)row[2];
            acc += k3 * (int)row[3];
            *dst = (uint8_t)clip8(acc);
          }
        } else {
          const uint8_t* row = inb + x0;
          for (int r = 0; r < nrows; ++r, row += w, dst += inter_stride) {
            int acc = 1 << (kPrecisionBits - 1);
            for (int t = 0; t < c; ++t) acc += k[t] * (int)row[t];
            *dst = (uint8_t)clip8(acc);
          }
        }
      }
      have_r0 = r0;
      have_r1 = r1;
      __syncthreads();
    }

    // (3) vertical pass: 4 adjacent output pixels per thread, one 32-bit load per tap; results staged linearly
    const int nby = y1 - y0;
    if (ty < lanes_y) {
      for (int yy = ty; yy < nby; yy += lanes_y) {
        const int y = y0 + yy;
        const int tb = tab_y[y];
        const int c = tb >> 20;
        const int* k = ky + y * ksy;
        const uint8_t* col = inter + ((tb & 0xFFFFF) - have_r0) * inter_stride + 4 * tx;
        int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0, a3 = a0;
        for (int t = 0; t < c; ++t, col += inter_stride) {
          const int kv = k[t];
          const uint32_t v = *reinterpret_cast<const uint32_t*>(col);
          a0 += kv * (int)(v & 255u);
          a1 += kv * (int)((v >> 8) & 255u);
          a2 += kv * (int)((v >> 16) & 255u);
          a3 += kv * (int)(v >> 24);
        }
        uint8_t* s = stage + yy * R + 4 * tx;
        const int x = 4 * tx;
        s[0] = (uint8_t)clip8(a0);
        if (x + 1 < R) s[1] = (uint8_t)clip8(a1);
        if (x + 2 < R) s[2] = (uint8_t)clip8(a2);
        if (x + 3 < R) s[3] = (uint8_t)clip8(a3);
      }
    }
    __syncthreads();
    // (4) write-out of the band's rows (contiguous in the output)
    write_out(p, stage, lut, roi, y0, nby * R);
    __syncthreads();
    y0 = y1;
  }
}

// shared memory of one launch for ROIs bounded by (max_h, max_w); returns 0 if even the direct mode does not fit
int plan_smem(int max_h, int max_w, int R, int* direct) {
  const int ksx = resample_ksize(max_w, R), ksy = resample_ksize(max_h, R);
  const long long inter_stride = (R + 3) & ~3;
  const long long fixed = 3072ll + 8ll * R + 4ll * R * (ksx + ksy) + 16 + 16 + ((kOutBand * R + 4 + 15) & ~15);
  const long long need_min = fixed + (long long)(ksy + 1) * (max_w + inter_stride) + 160;
  const long long whole = fixed + (long long)max_h * (max_w + inter_stride) + 160;
  *direct = 0;
  if (need_min <= kSmemMax) {
    long long s = whole < kSmemTarget ? whole : kSmemTarget;
    if (s < need_min) s = need_min;
    return (int)((s + 1023) & ~1023ll) > kSmemMax ? kSmemMax : (int)((s + 1023) & ~1023ll);
  }
  const long long need_direct = fixed + (long long)(ksy + 1) * inter_stride + 160;
  if (need_direct > kSmemMax) return 0;
  *direct = 1;
  return (int)((need_direct + 1023) & ~1023ll) > kSmemMax ? kSmemMax : (int)((need_direct + 1023) & ~1023ll);
}

}  // namespace
}  // namespace ifcb

extern "C" int ifcb_preprocess(const uint8_t* d_packed, int64_t packed_bytes,
                               const int64_t* d_offsets, const int32_t* d_h, const int32_t* d_w,
                               int n, int max_h, int max_w, int R,
                               const float* h_mean, const float* h_std,
                               int out_mode, void* d_out, int pass_rule, int32_t* d_status, void* stream) {
  using namespace ifcb;
  IFCB_ARG_CHECK(n >= 0, "ifcb_preprocess: n < 0");
  if (n == 0) return 0;
  IFCB_ARG_CHECK(d_packed && d_offsets && d_h && d_w && d_out, "ifcb_preprocess: null pointer");
  IFCB_ARG_CHECK(R >= 1 && R <= 512, "ifcb_preprocess: R=%d out of range [1,512]", R);
  IFCB_ARG_CHECK(max_h >= 1 && max_w >= 1, "ifcb_preprocess: max_h/max_w must be >= 1");
  IFCB_ARG_CHECK(max_h < (1 << 20) && max_w < (1 << 20), "ifcb_preprocess: max_h/max_w out of range");
  int direct = 0;
  const int smem = plan_smem(max_h, max_w, R, &direct);
  IFCB_ARG_CHECK(smem > 0, "ifcb_preprocess: ROI bound %dx%d -> %d needs more than %d B of shared memory even without input staging",
                 max_h, max_w, R, kSmemMax);
  IFCB_ARG_CHECK(out_mode >= 0 && out_mode <= 2, "ifcb_preprocess: bad out_mode %d", out_mode);
  IFCB_ARG_CHECK(pass_rule == IFCB_PASS_PILLOW12 || pass_rule == IFCB_PASS_HV,
                 "ifcb_preprocess: bad pass_rule %d", pass_rule);
  IFCB_ARG_CHECK((reinterpret_cast<uintptr_t>(d_packed) & 15) == 0,
                 "ifcb_preprocess: d_packed must be 16-byte aligned");
  PreParams p;
  p.packed = d_packed;
  p.packed_bytes = packed_bytes;
  p.offsets = reinterpret_cast<const long long*>(d_offsets);
  p.hs = d_h;
  p.ws = d_w;
  p.n = n;
  p.R = R;
  p.out_mode = out_mode;
  p.pass_rule = pass_rule;
  p.has_norm = (h_mean && h_std) ? 1 : 0;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = p.has_norm ? h_mean[c] : 0.f;
    p.stdv[c] = p.has_norm ? h_std[c] : 1.f;
  }
  p.out = d_out;
  p.smem_bytes = smem;
  p.direct = direct;
  p.status = d_status;
  // the opt-in shared-memory limit is a per-device function attribute
  static bool attr_set[64] = {};
  int dev = 0;
  IFCB_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    IFCB_CUDA_CHECK(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  preprocess_kernel<<<n, kThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
