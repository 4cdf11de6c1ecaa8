// TRAIN-step kernels around the tensor-core GEMMs (sm_100a): batch-norm statistics / apply /
// backward, pooling forward-with-index and backward, gradient dilation for strided convs,
// the classifier head with cross-entropy, Adam and the weight repack that feeds the tcgen05
// forward / data-gradient kernels.
//
// Replaces what torch autograd + torch.optim.Adam run for NeustonModel.training_step
// (reference neuston_models.py:63-86: train-mode forward of the torchvision graph,
// CrossEntropyLoss (+0.4 * aux), loss.backward(), Adam.step()).
//
// All of these are HBM-bound streaming kernels: one thread = one pixel x 8 channels (16 bytes),
// NHWC 16-bit views with an optional physical zero border (include/ifcb_b200.h: ifcb_view).
#include "layers.cuh"
#include <cstdlib>

namespace ifcb {
namespace {

struct DV {   // device copy of an ifcb_view
  uint16_t* p;
  int ld, C, H, W, ph, pw;
  unsigned long long magic_w, magic_h;   // fast_div magics of W and H
};

inline DV dv(const ifcb_view* v) {
  DV d;
  d.p = reinterpret_cast<uint16_t*>(v->d);
  d.ld = v->ld; d.C = v->C; d.H = v->H; d.W = v->W; d.ph = v->pad_h; d.pw = v->pad_w;
  d.magic_w = div_magic(v->W);
  d.magic_h = div_magic(v->H);
  return d;
}

__device__ __forceinline__ long long pix_off(const DV& v, int n, int h, int w) {
  return (((long long)n * (v.H + 2 * v.ph) + h + v.ph) * (v.W + 2 * v.pw) + w + v.pw) * v.ld;
}

// logical pixel index m = (n*H + h)*W + w (< 2^31)  ->  element offset of that pixel in the view
__device__ __forceinline__ long long pix_off_m(const DV& v, long long m) {
  if ((v.ph | v.pw) == 0) return m * v.ld;
  const uint32_t t = fast_div((uint32_t)m, v.magic_w);
  const int w = (int)((uint32_t)m - t * (uint32_t)v.W);
  const uint32_t n = fast_div(t, v.magic_h);
  const int h = (int)(t - n * (uint32_t)v.H);
  return pix_off(v, (int)n, h, w);
}

__device__ __forceinline__ void load8(const uint16_t* p, int fp16, float* f) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = unpack_act2(u[j], fp16);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

__device__ __forceinline__ void store8(uint16_t* p, int fp16, const float* f) {
  uint4 o;
  o.x = pack_act2(f[0], f[1], fp16);
  o.y = pack_act2(f[2], f[3], fp16);
  o.z = pack_act2(f[4], f[5], fp16);
  o.w = pack_act2(f[6], f[7], fp16);
  *reinterpret_cast<uint4*>(p) = o;
}

bool view_ok(const ifcb_view* v) {
  return v && v->d && v->C > 0 && v->C % 8 == 0 && v->ld >= v->C && v->ld % 8 == 0 && v->H > 0 && v->W > 0 &&
         v->pad_h >= 0 && v->pad_w >= 0 && (reinterpret_cast<uintptr_t>(v->d) & 15) == 0;
}

int grid_for(long long total, int threads, int cap_per_sm = 16) {
  long long g = (total + threads - 1) / threads;
  const long long cap = (long long)sm_count() * cap_per_sm;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

// ------------------------------------------------------------------------------------------
// Per-channel reductions over all pixels of a view.  Block = rows x (C/8) threads; every thread
// keeps fp32 partials of its 8 channels over its pixel rows (grid-stride), the block reduces
// over rows in shared memory and adds into float64 accumulators (order-insensitive to fp32).
//   MODE 0 (bn_stats):      acc[c] += z, acc[C + c] += z*z
//   MODE 1 (bn_bwd_reduce): dy' = relu-masked dy; acc[c] += dy', acc[C + c] += dy' * (z - mean) * invstd
// ------------------------------------------------------------------------------------------
// BN output before the activation, shared by forward and backward so that the ReLU mask the
// backward pass recomputes from z is bit-identical to the forward's:
//   y = fma(z - mean, S, beta),  S = gamma * invstd      (z - mean first: exact near the mean)
// bn_params8 loads the three per-channel vectors a thread keeps in registers.
__device__ __forceinline__ void ldg8(const float* __restrict__ p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__device__ __forceinline__ void bn_params8(const float* __restrict__ mean, const float* __restrict__ invstd,
                                           const float* __restrict__ gamma, const float* __restrict__ beta, int c0, float* mu, float* S,
                                           float* Bt) {
  float is[8], ga[8];
  ldg8(mean + c0, mu);
  ldg8(invstd + c0, is);
  ldg8(gamma + c0, ga);
  ldg8(beta + c0, Bt);
#pragma unroll
  for (int j = 0; j < 8; ++j) S[j] = __fmul_rn(ga[j], is[j]);
}

__device__ __forceinline__ float bn_y(float z, float mu, float S, float b) { return fmaf(__fsub_rn(z, mu), S, b); }

// mask modes of the backward kernels
enum { kMaskNone = 0, kMaskFromZ = 1, kMaskFromA = 2 };

constexpr int kUnroll = 4;   // pixel rows in flight per thread in the streaming kernels

__device__ __forceinline__ uint4 ld16(const uint16_t* p) { return *reinterpret_cast<const uint4*>(p); }

__device__ __forceinline__ void cvt8(const uint4& v, int fp16, float* f) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = unpack_act2(u[j], fp16);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

// What the LAST block of a reduction does with the complete float64 sums (saves one tiny launch per layer and pass):
//   MODE 0 (forward): mean / invstd, running-stat update as torch.nn.BatchNorm2d in train mode (momentum, UNBIASED
//                     variance into running_var)
//   MODE 1 (backward): per-channel coefficients of the elementwise pass, dz = A * dy' + Bz * z + D with
//                     A = gamma*invstd, Bz = -A*invstd*s2/M, D = A*(mean*invstd*s2/M - s1/M); dgamma += s2, dbeta += s1
// and clears the accumulators for the next user.
struct FinalizeArgs {
  double count;            // pixels per channel
  float eps, momentum;
  float* mean_out;         // MODE 0 outputs
  float* invstd_out;
  float* run_mean;
  float* run_var;
  float* coef;             // MODE 1 outputs: [A | Bz | D], 3*C floats
  float* dgamma;
  float* dbeta;
  unsigned int* counter;   // blocks finished so far (zero on entry, zero again on exit)
  float* det_part;         // deterministic mode: per-block partial sums [channel block][pixel split][2 * channels of the block],
                           // stored (no atomics) and summed by the finalising block in pixel-split order
};

// MODE 0: sums of z and z*z.   MODE 1: sums of dy' and dy' * (z - mean) (scaled by invstd at the end).
template <int MODE, int MASK>
__global__ void __launch_bounds__(256, 2) channel_reduce_kernel(DV z, DV dy, DV a, const float* __restrict__ mean,
                                                                const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, long long M, int rows, int fp16,
                                                                double* __restrict__ acc, const FinalizeArgs fin, int c8b) {
  // grid = (pixel splits, channel blocks): a block owns c8b groups of 8 channels (<= 64 channels = one 128-byte line per pixel)
  // and `rows` pixel rows per sweep.  Small channel blocks keep the float64 atomics per block at 2 * 8 * c8b whatever C is, so
  // small tensors can be spread over a full wave of blocks (the first version gave every block ALL channels: 2 * C atomics per
  // block forced 49-196 blocks on the 7x7 / 14x14 layers -- 22-45 us for tensors that stream in 5 us).
  __shared__ float red[256 * 16];
  __shared__ int s_last;
  const int c8n = c8b;
  const int tid = threadIdx.x;
  const int row = tid / c8n, c8 = blockIdx.y * c8b + (tid - row * c8n);
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (row < rows) {
    float mu[8], S[8], Bt[8];
    if (MODE == 1) {
      if (MASK == kMaskFromZ) bn_params8(mean, invstd, gamma, beta, c8 * 8, mu, S, Bt);
      else ldg8(mean + c8 * 8, mu);
    }
    const long long stride = (long long)gridDim.x * rows;
    for (long long m0 = (long long)blockIdx.x * rows + row; m0 < M; m0 += stride * kUnroll) {
      uint4 zr[kUnroll], gr[kUnroll], ar[kUnroll];
      // all loads of the kUnroll rows first (independent), then the arithmetic
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long m = m0 + u * stride;
        if (m < M) {
          zr[u] = ld16(z.p + pix_off_m(z, m) + c8 * 8);
          if (MODE == 1) {
            gr[u] = ld16(dy.p + pix_off_m(dy, m) + c8 * 8);
            if (MASK == kMaskFromA) ar[u] = ld16(a.p + pix_off_m(a, m) + c8 * 8);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (m0 + u * stride < M) {
          float zv[8];
          cvt8(zr[u], fp16, zv);
          if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              s1[j] += zv[j];
              s2[j] = fmaf(zv[j], zv[j], s2[j]);
            }
          } else {
            float g[8];
            cvt8(gr[u], fp16, g);
            if (MASK == kMaskFromZ) {
#pragma unroll
              for (int j = 0; j < 8; ++j) g[j] = bn_y(zv[j], mu[j], S[j], Bt[j]) > 0.f ? g[j] : 0.f;
            } else if (MASK == kMaskFromA) {
              float av[8];
              cvt8(ar[u], fp16, av);
#pragma unroll
              for (int j = 0; j < 8; ++j) g[j] = av[j] > 0.f ? g[j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              s1[j] += g[j];
              s2[j] = fmaf(g[j], zv[j] - mu[j], s2[j]);
            }
          }
        }
      }
    }
    if (MODE == 1) {
      float is[8];
      ldg8(invstd + c8 * 8, is);
#pragma unroll
      for (int j = 0; j < 8; ++j) s2[j] *= is[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[tid * 16 + j] = s1[j];
    red[tid * 16 + 8 + j] = s2[j];
  }
  __syncthreads();
  // thread t < 2 * (channels of this block) sums one (quantity, channel) over the rows
  const int C = z.C, cb = 8 * c8b, c_lo = blockIdx.y * cb;
  for (int t = tid; t < 2 * cb; t += 256) {
    const int which = t / cb, cl = t - which * cb;
    const int cc8 = cl >> 3, j = cl & 7;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += red[(r * c8n + cc8) * 16 + which * 8 + j];
    if (fin.det_part) fin.det_part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (2 * cb) + t] = s;
    else atomicAdd(acc + which * C + c_lo + cl, (double)s);
  }
  // ---- last block of this channel block: finalize its channels (threadFenceReduction pattern) ----
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(fin.counter + blockIdx.y, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  for (int c = c_lo + tid; c < c_lo + cb; c += 256) {
    double s1, s2;
    if (fin.det_part) {                     // fixed-order float64 sum of the blocks' partials
      s1 = s2 = 0.0;
      const float* pp = fin.det_part + (size_t)blockIdx.y * gridDim.x * (2 * cb) + (c - c_lo);
      for (unsigned b = 0; b < gridDim.x; ++b) {
        s1 += (double)__ldcg(pp + (size_t)b * (2 * cb));
        s2 += (double)__ldcg(pp + (size_t)b * (2 * cb) + cb);
      }
    } else {
      s1 = __ldcg(acc + c);
      s2 = __ldcg(acc + C + c);
    }
    if (MODE == 0) {
      const double mu = s1 / fin.count;
      double var = s2 / fin.count - mu * mu;
      if (var < 0.0) var = 0.0;
      fin.mean_out[c] = (float)mu;
      fin.invstd_out[c] = (float)(1.0 / sqrt(var + (double)fin.eps));
      if (fin.run_mean) fin.run_mean[c] = (1.f - fin.momentum) * fin.run_mean[c] + fin.momentum * (float)mu;
      if (fin.run_var) {
        const double unb = fin.count > 1.0 ? var * fin.count / (fin.count - 1.0) : var;
        fin.run_var[c] = (1.f - fin.momentum) * fin.run_var[c] + fin.momentum * (float)unb;
      }
    } else {
      const double inv_m = 1.0 / fin.count;
      const double is = (double)invstd[c], mu = (double)mean[c];
      const double A = (double)gamma[c] * is;
      fin.coef[c] = (float)A;
      fin.coef[C + c] = (float)(-A * is * s2 * inv_m);
      fin.coef[2 * C + c] = (float)(A * (mu * is * s2 * inv_m - s1 * inv_m));
      if (fin.dbeta) fin.dbeta[c] += (float)s1;
      if (fin.dgamma) fin.dgamma[c] += (float)s2;
    }
    acc[c] = 0.0;
    acc[C + c] = 0.0;
  }
  if (tid == 0) fin.counter[blockIdx.y] = 0u;
}

// a = act(bn_y(z) (+ residual)).  Block = rows x (C/8) threads: a thread keeps the parameters of its 8
// channels in registers and streams pixel rows, kUnroll rows in flight.
template <bool HAS_RES>
__global__ void __launch_bounds__(256, 3) bn_apply_kernel(DV z, DV out, DV res, const float* __restrict__ mean,
                                                          const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, int relu, long long M, int rows, int fp16) {
  const int c8n = z.C >> 3;
  const int row = threadIdx.x / c8n, c8 = threadIdx.x - row * c8n;
  if (row >= rows) return;
  float mu[8], S[8], Bt[8];
  bn_params8(mean, invstd, gamma, beta, c8 * 8, mu, S, Bt);
  const long long stride = (long long)gridDim.x * rows;
  for (long long m0 = (long long)blockIdx.x * rows + row; m0 < M; m0 += stride * kUnroll) {
    uint4 zr[kUnroll], rr[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long m = m0 + u * stride;
      if (m < M) {
        zr[u] = ld16(z.p + pix_off_m(z, m) + c8 * 8);
        if (HAS_RES) rr[u] = ld16(res.p + pix_off_m(res, m) + c8 * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long m = m0 + u * stride;
      if (m < M) {
        float v[8], r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cvt8(zr[u], fp16, v);
        if (HAS_RES) cvt8(rr[u], fp16, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float y = bn_y(v[j], mu[j], S[j], Bt[j]) + r[j];
          v[j] = relu ? fmaxf(y, 0.f) : y;
        }
        store8(out.p + pix_off_m(out, m) + c8 * 8, fp16, v);
      }
    }
  }
}

// The same pass when the batch statistics arrive as float64 sums gathered by the convolution's epilogue (ifcb_conv_desc.d_stats):
// every thread turns the sums of ITS 8 channels into mean / invstd (a handful of float64 operations, redundant across blocks but
// free next to the streaming), block 0 also publishes mean / invstd for the backward pass and updates the running statistics
// exactly as the last block of channel_reduce_kernel<0> does.  No pass over z for the statistics, no extra launch.
template <bool HAS_RES>
__global__ void __launch_bounds__(256, 3) bn_apply_sums_kernel(DV z, DV out, DV res, const double* __restrict__ sums, double count, float eps,
                                                               float momentum, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                                               float* __restrict__ run_mean, float* __restrict__ run_var,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                                                               long long M, int rows, int fp16) {
  const int C = z.C, c8n = C >> 3;
  const int row = threadIdx.x / c8n, c8 = threadIdx.x - row * c8n;
  if (row >= rows) return;
  float mu[8], S[8], Bt[8];
  {
    float ga[8];
    ldg8(gamma + c8 * 8, ga);
    ldg8(beta + c8 * 8, Bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
      const double m = sums[c] / count;
      double var = sums[C + c] / count - m * m;
      if (var < 0.0) var = 0.0;
      const float is = (float)(1.0 / sqrt(var + (double)eps));
      mu[j] = (float)m;
      S[j] = __fmul_rn(ga[j], is);
      if (blockIdx.x == 0 && row == 0) {
        mean_out[c] = (float)m;
        invstd_out[c] = is;
        if (run_mean) run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)m;
        if (run_var) {
          const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
          run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unb;
        }
      }
    }
  }
  const long long stride = (long long)gridDim.x * rows;
  for (long long m0 = (long long)blockIdx.x * rows + row; m0 < M; m0 += stride * kUnroll) {
    uint4 zr[kUnroll], rr[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long m = m0 + u * stride;
      if (m < M) {
        zr[u] = ld16(z.p + pix_off_m(z, m) + c8 * 8);
        if (HAS_RES) rr[u] = ld16(res.p + pix_off_m(res, m) + c8 * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long m = m0 + u * stride;
      if (m < M) {
        float v[8], r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cvt8(zr[u], fp16, v);
        if (HAS_RES) cvt8(rr[u], fp16, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float y = bn_y(v[j], mu[j], S[j], Bt[j]) + r[j];
          v[j] = relu ? fmaxf(y, 0.f) : y;
        }
        store8(out.p + pix_off_m(out, m) + c8 * 8, fp16, v);
      }
    }
  }
}

// dz = A * dy' + Bz * z + D with dy' = relu-masked dy; optionally routes dy' to the residual branch
// (RES 1: written, 2: accumulated).  dz may alias dy (in place).  Same thread mapping as bn_apply_kernel,
// two rows in flight (the six per-channel vectors already take 48 registers).
template <int MASK, int RES, bool ACC = false>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(DV dy, DV a, DV z, DV dz, DV dres, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ coef,
                                                              long long M, int rows, int fp16) {
  constexpr int U = 4;
  const int C = z.C, c8n = C >> 3;
  const int row = threadIdx.x / c8n, c8 = threadIdx.x - row * c8n;
  if (row >= rows) return;
  float mu[8], S[8], Bt[8], A[8], Bz[8], D[8];
  if (MASK == kMaskFromZ) bn_params8(mean, invstd, gamma, beta, c8 * 8, mu, S, Bt);
  ldg8(coef + c8 * 8, A);
  ldg8(coef + C + c8 * 8, Bz);
  ldg8(coef + 2 * C + c8 * 8, D);
  const long long stride = (long long)gridDim.x * rows;
  for (long long m0 = (long long)blockIdx.x * rows + row; m0 < M; m0 += stride * U) {
    uint4 gr[U], zr[U], ar[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = m0 + u * stride;
      if (m < M) {
        gr[u] = ld16(dy.p + pix_off_m(dy, m) + c8 * 8);
        zr[u] = ld16(z.p + pix_off_m(z, m) + c8 * 8);
        if (MASK == kMaskFromA) ar[u] = ld16(a.p + pix_off_m(a, m) + c8 * 8);
        if (RES == 2) rr[u] = ld16(dres.p + pix_off_m(dres, m) + c8 * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = m0 + u * stride;
      if (m < M) {
        float g[8], zv[8];
        cvt8(gr[u], fp16, g);
        cvt8(zr[u], fp16, zv);
        if (MASK == kMaskFromZ) {
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = bn_y(zv[j], mu[j], S[j], Bt[j]) > 0.f ? g[j] : 0.f;
        } else if (MASK == kMaskFromA) {
          float av[8];
          cvt8(ar[u], fp16, av);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = av[j] > 0.f ? g[j] : 0.f;
        }
        if (RES) {
          float r[8];
          if (RES == 2) {
            cvt8(rr[u], fp16, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] += g[j];
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = g[j];
          }
          store8(dres.p + pix_off_m(dres, m) + c8 * 8, fp16, r);
        }
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], g[j], fmaf(Bz[j], zv[j], D[j]));
        if (ACC) {                       // dz += ... (the target already holds other consumers' contributions)
          float old[8];
          load8(dz.p + pix_off_m(dz, m) + c8 * 8, fp16, old);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += old[j];
        }
        store8(dz.p + pix_off_m(dz, m) + c8 * 8, fp16, o);
      }
    }
  }
}

// Backward of (conv + bias -> [ReLU]) for the families without BatchNorm (AlexNet, VGG, SqueezeNet; torchvision alexnet.py /
// vgg.py / squeezenet.py reached from NeustonModel.training_step, neuston_models.py:80-86): dz = dy * [a > 0] written (dz may
// alias dy), dbias[c] += sum over pixels of dz.  One pass; the same (pixel splits x channel blocks) grid and float64 / ordered
// partial-sum machinery as channel_reduce_kernel.
template <bool RELU>
__global__ void __launch_bounds__(256, 2) bias_relu_bwd_kernel(DV dy, DV a, DV dz, long long M, int rows, int fp16, double* __restrict__ acc,
                                                               float* __restrict__ dbias, unsigned int* __restrict__ counter,
                                                               float* __restrict__ det_part, int c8b) {
  __shared__ float red[256 * 8];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int row = tid / c8b, c8 = blockIdx.y * c8b + (tid - row * c8b);
  float s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = 0.f;
  if (row < rows) {
    const long long stride = (long long)gridDim.x * rows;
    for (long long m0 = (long long)blockIdx.x * rows + row; m0 < M; m0 += stride * kUnroll) {
      uint4 gr[kUnroll], ar[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long m = m0 + u * stride;
        if (m < M) {
          gr[u] = ld16(dy.p + pix_off_m(dy, m) + c8 * 8);
          if (RELU) ar[u] = ld16(a.p + pix_off_m(a, m) + c8 * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long m = m0 + u * stride;
        if (m < M) {
          float g[8];
          cvt8(gr[u], fp16, g);
          if (RELU) {
            float av[8];
            cvt8(ar[u], fp16, av);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = av[j] > 0.f ? g[j] : 0.f;
            store8(dz.p + pix_off_m(dz, m) + c8 * 8, fp16, g);
          } else if (dz.p != dy.p) {
            store8(dz.p + pix_off_m(dz, m) + c8 * 8, fp16, g);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) s1[j] += g[j];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[tid * 8 + j] = s1[j];
  __syncthreads();
  const int C = dy.C, cb = 8 * c8b, c_lo = blockIdx.y * cb;
  for (int t = tid; t < cb; t += 256) {
    const int cc8 = t >> 3, j = t & 7;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += red[(r * c8b + cc8) * 8 + j];
    if (det_part) det_part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * cb + t] = s;
    else atomicAdd(acc + c_lo + t, (double)s);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(counter + blockIdx.y, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  for (int c = c_lo + tid; c < c_lo + cb && c < C; c += 256) {
    double s = 0.0;
    if (det_part) {
      const float* pp = det_part + (size_t)blockIdx.y * gridDim.x * cb + (c - c_lo);
      for (unsigned b = 0; b < gridDim.x; ++b) s += (double)__ldcg(pp + (size_t)b * cb);
    } else {
      s = __ldcg(acc + c);
      acc[c] = 0.0;
    }
    dbias[c] += (float)s;
  }
  if (tid == 0) counter[blockIdx.y] = 0u;
}

// y = x * scale, element-wise, scale float32 in the logical [batch, H, W, C] order (dropout masks of the classifier stacks:
// alexnet.py / vgg.py / squeezenet.py nn.Dropout in train mode; the backward pass applies the same scale to the gradient)
__global__ void __launch_bounds__(256) scale_elems_kernel(DV x, DV y, const float* __restrict__ scale, uint32_t total,
                                                          unsigned long long magic_c8, int fp16) {
  const uint32_t c8n = (uint32_t)(x.C >> 3);
  for (uint32_t t = blockIdx.x * 256u + threadIdx.x; t < total; t += gridDim.x * 256u) {
    const uint32_t m = fast_div(t, magic_c8);
    const int c8 = (int)(t - m * c8n);
    float v[8], sc[8];
    load8(x.p + pix_off_m(x, m) + c8 * 8, fp16, v);
    ldg8(scale + (long long)m * x.C + c8 * 8, sc);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= sc[j];
    store8(y.p + pix_off_m(y, m) + c8 * 8, fp16, v);
  }
}

// ------------------------------------------------------------------------------------------
// Pooling.
// ------------------------------------------------------------------------------------------
// max pool forward that also records the winning tap (first maximum in window scan order, as
// torch.nn.functional.max_pool2d) for the backward pass.  idx: uint8 [B, P, Q, C].
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(DV x, DV y, uint8_t* __restrict__ idx, int k, int stride, int pad,
                                                          int P, int Q, uint32_t total, unsigned long long magic_c8, int fp16) {
  const uint32_t c8n = (uint32_t)(x.C >> 3);
  for (uint32_t t = blockIdx.x * 256u + threadIdx.x; t < total; t += gridDim.x * 256u) {
    const uint32_t m = fast_div(t, magic_c8);
    const int c8 = (int)(t - m * c8n);
    const uint32_t t2 = fast_div(m, y.magic_w);
    const int oq = (int)(m - t2 * (uint32_t)Q);
    const uint32_t nn = fast_div(t2, y.magic_h);
    const int op = (int)(t2 - nn * (uint32_t)P), n = (int)nn;
    const int h0 = op * stride - pad, w0 = oq * stride - pad;
    float best[8];
    int bi[8];
    bool first = true;
    for (int r = 0; r < k; ++r) {
      const int hh = h0 + r;
      if (hh < 0 || hh >= x.H) continue;
      for (int s = 0; s < k; ++s) {
        const int ww = w0 + s;
        if (ww < 0 || ww >= x.W) continue;
        float v[8];
        load8(x.p + pix_off(x, n, hh, ww) + c8 * 8, fp16, v);
        const int tap = r * k + s;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (first || v[j] > best[j] || v[j] != v[j]) {
            best[j] = v[j];
            bi[j] = tap;
          }
        }
        first = false;
      }
    }
    store8(y.p + pix_off(y, n, op, oq) + c8 * 8, fp16, best);
    uint2 pk;
    pk.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
    pk.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
    *reinterpret_cast<uint2*>(idx + (long long)m * x.C + c8 * 8) = pk;
  }
}

// the same for 3x3 windows (every max pool of ResNet / Inception-v3), in PACKED 16-bit arithmetic: these kernels are bound by
// instruction issue, not by memory (the float version spent ~360 instructions per 16 output bytes and ran at 2.4 TB/s).  The nine
// 16-byte loads are issued together; per filter tap and pair of channels: one packed compare mask (v > best, false when unordered),
// one packed max (NaN propagates) and one LOP3 that moves the tap number into the lanes that won -- strict '>' in window-scan
// order keeps the FIRST maximum, as torch.nn.functional.max_pool2d.  Values are never converted: the output is bit-exact.
template <bool FP16>
__device__ __forceinline__ uint32_t gt2_mask(uint32_t a, uint32_t b) {
  if (FP16) return __hgt2_mask(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
}
template <bool FP16>
__device__ __forceinline__ uint32_t max2_nan(uint32_t a, uint32_t b) {
  if (FP16) {
    const __half2 r = __hmax2_nan(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const __nv_bfloat162 r = __hmax2_nan(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

template <bool FP16>
__global__ void __launch_bounds__(256) maxpool3_fwd_kernel(DV x, DV y, uint8_t* __restrict__ idx, int stride, int pad, int P, int Q,
                                                           uint32_t total, unsigned long long magic_c8) {
  const uint32_t c8n = (uint32_t)(x.C >> 3);
  const long long row_pitch = (long long)(x.W + 2 * x.pw) * x.ld;
  const uint32_t neg_inf = FP16 ? 0xfc00fc00u : 0xff80ff80u;
  for (uint32_t t = blockIdx.x * 256u + threadIdx.x; t < total; t += gridDim.x * 256u) {
    const uint32_t m = fast_div(t, magic_c8);
    const int c8 = (int)(t - m * c8n);
    const uint32_t t2 = fast_div(m, y.magic_w);
    const int oq = (int)(m - t2 * (uint32_t)Q);
    const uint32_t nn = fast_div(t2, y.magic_h);
    const int op = (int)(t2 - nn * (uint32_t)P), n = (int)nn;
    const int h0 = op * stride - pad, w0 = oq * stride - pad;
    const uint16_t* base = x.p + pix_off(x, n, h0, w0) + c8 * 8;      // dereferenced only where the tap is inside the image
    uint4 v[9];
    bool ok[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        ok[r * 3 + s2] = (unsigned)(h0 + r) < (unsigned)x.H && (unsigned)(w0 + s2) < (unsigned)x.W;
        if (ok[r * 3 + s2]) v[r * 3 + s2] = ld16(base + r * row_pitch + s2 * x.ld);
      }
    uint32_t best[4] = {neg_inf, neg_inf, neg_inf, neg_inf};
    uint32_t bi[4] = {0u, 0u, 0u, 0u};                 // winning tap per 16-bit lane
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      if (ok[tap]) {
        const uint32_t u[4] = {v[tap].x, v[tap].y, v[tap].z, v[tap].w};
        const uint32_t tapc = (uint32_t)tap * 0x00010001u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t won = gt2_mask<FP16>(u[j], best[j]);
          best[j] = max2_nan<FP16>(best[j], u[j]);
          bi[j] = (bi[j] & ~won) | (tapc & won);
        }
      }
    }
    *reinterpret_cast<uint4*>(y.p + pix_off(y, n, op, oq) + c8 * 8) = make_uint4(best[0], best[1], best[2], best[3]);
    uint2 pk;                                          // 16-bit lanes -> bytes
    pk.x = __byte_perm(bi[0], bi[1], 0x6420);
    pk.y = __byte_perm(bi[2], bi[3], 0x6420);
    *reinterpret_cast<uint2*>(idx + (long long)m * x.C + c8 * 8) = pk;
  }
}

// gather form of the max-pool / avg-pool backward: one thread = one INPUT pixel x 8 channels,
// sums the gradients of the output windows that cover it (deterministic, no atomics).
template <bool AVG>
__global__ void __launch_bounds__(256) pool_bwd_kernel(DV dy, const uint8_t* __restrict__ idx, DV dx, int accumulate, int k,
                                                       int stride, int pad, int P, int Q, uint32_t total, unsigned long long magic_c8,
                                                       int fp16) {
  const uint32_t c8n = (uint32_t)(dx.C >> 3);
  const float inv = 1.f / (float)(k * k);
  for (uint32_t t = blockIdx.x * 256u + threadIdx.x; t < total; t += gridDim.x * 256u) {
    const uint32_t m = fast_div(t, magic_c8);
    const int c8 = (int)(t - m * c8n);
    const uint32_t t2 = fast_div(m, dx.magic_w);
    const int w = (int)(m - t2 * (uint32_t)dx.W);
    const uint32_t nn = fast_div(t2, dx.magic_h);
    const int h = (int)(t2 - nn * (uint32_t)dx.H), n = (int)nn;
    float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // windows op with op*stride - pad <= h <= op*stride - pad + k - 1
    int p_lo = h + pad - k + 1;
    p_lo = p_lo <= 0 ? 0 : (p_lo + stride - 1) / stride;
    int p_hi = (h + pad) / stride;
    if (p_hi > P - 1) p_hi = P - 1;
    int q_lo = w + pad - k + 1;
    q_lo = q_lo <= 0 ? 0 : (q_lo + stride - 1) / stride;
    int q_hi = (w + pad) / stride;
    if (q_hi > Q - 1) q_hi = Q - 1;
    for (int op = p_lo; op <= p_hi; ++op) {
      const int r = h + pad - op * stride;
      for (int oq = q_lo; oq <= q_hi; ++oq) {
        const int s = w + pad - oq * stride;
        float v[8];
        load8(dy.p + pix_off(dy, n, op, oq) + c8 * 8, fp16, v);
        if (AVG) {
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += v[j];
        } else {
          const uint2 pk = *reinterpret_cast<const uint2*>(idx + (((long long)n * P + op) * Q + oq) * dx.C + c8 * 8);
          const int tap = r * k + s;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int w8 = (int)(((j < 4 ? pk.x : pk.y) >> (8 * (j & 3))) & 0xffu);
            if (w8 == tap) g[j] += v[j];
          }
        }
      }
    }
    if (AVG) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= inv;
    }
    uint16_t* dst = dx.p + pix_off(dx, n, h, w) + c8 * 8;
    if (accumulate) {
      float o[8];
      load8(dst, fp16, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += o[j];
    }
    store8(dst, fp16, g);
  }
}

// max-pool backward, 3x3 window / stride 2 (every max pool of ResNet and Inception-v3): an input pixel lies in at most 2 x 2
// windows, so the four (gradient, winner-index) loads are issued together and combined afterwards -- the generic gather above
// walks the windows one by one.  Issue-bound like the forward kernel: the winner test is one SIMD byte compare per four channels
// (__vcmpeq4 against the tap this pixel is in that window), the byte masks are widened to 16-bit lanes (PRMT) and ANDed onto
// the packed gradients, and only then converted and summed in fp32 -- same sum, same order (rows, then columns) as the generic
// kernel, ~2x fewer instructions.
template <bool FP16>
__global__ void __launch_bounds__(256) maxpool3s2_bwd_kernel(DV dy, const uint8_t* __restrict__ idx, DV dx, int accumulate, int pad, int P,
                                                             int Q, uint32_t total, unsigned long long magic_c8) {
  const uint32_t c8n = (uint32_t)(dx.C >> 3);
  for (uint32_t t = blockIdx.x * 256u + threadIdx.x; t < total; t += gridDim.x * 256u) {
    const uint32_t m = fast_div(t, magic_c8);
    const int c8 = (int)(t - m * c8n);
    const uint32_t t2 = fast_div(m, dx.magic_w);
    const int w = (int)(m - t2 * (uint32_t)dx.W);
    const uint32_t nn = fast_div(t2, dx.magic_h);
    const int h = (int)(t2 - nn * (uint32_t)dx.H), n = (int)nn;
    // windows op with 2*op - pad <= h <= 2*op - pad + 2  =>  op in [ceil((h + pad - 2) / 2), floor((h + pad) / 2)]
    const int hp = h + pad, wp = w + pad;
    const int p1 = hp >> 1, p0 = p1 - 1 + (hp & 1);          // hp even: {p1 - 1, p1}; hp odd: {p1} twice
    const int q1 = wp >> 1, q0 = q1 - 1 + (wp & 1);
    uint4 gv[4];
    uint2 iv[4];
    bool ok[4];
    uint32_t tapb[4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int op = a ? p1 : p0, oq = b ? q1 : q0;
        const int k = a * 2 + b;
        ok[k] = op >= 0 && op < P && oq >= 0 && oq < Q && (a == 1 || p0 != p1) && (b == 1 || q0 != q1);
        tapb[k] = (uint32_t)((hp - 2 * op) * 3 + (wp - 2 * oq)) * 0x01010101u;
        if (ok[k]) {
          gv[k] = ld16(dy.p + pix_off(dy, n, op, oq) + c8 * 8);
          iv[k] = *reinterpret_cast<const uint2*>(idx + (((long long)n * P + op) * Q + oq) * dx.C + c8 * 8);
        }
      }
    float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (ok[k]) {
        const uint32_t m0 = __vcmpeq4(iv[k].x, tapb[k]), m1 = __vcmpeq4(iv[k].y, tapb[k]);     // 0xff per channel this pixel won
        const uint32_t u[4] = {gv[k].x & __byte_perm(m0, 0, 0x1100), gv[k].y & __byte_perm(m0, 0, 0x3322),
                               gv[k].z & __byte_perm(m1, 0, 0x1100), gv[k].w & __byte_perm(m1, 0, 0x3322)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_act2(u[j], FP16 ? 1 : 0);
          g[2 * j] += f.x;
          g[2 * j + 1] += f.y;
        }
      }
    }
    uint16_t* dst = dx.p + pix_off(dx, n, h, w) + c8 * 8;
    if (accumulate) {
      float o[8];
      load8(dst, FP16 ? 1 : 0, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += o[j];
    }
    store8(dst, FP16 ? 1 : 0, g);
  }
}

// 3x3 / stride 1 / pad 1 box mean (count_include_pad: divisor 9) -- the forward AND the backward of every avg pool inside the
// Inception blocks (the backward of this pool is the same box mean of the gradient).  All nine 16-byte loads in flight, fp32 sum in
// window-scan order (the generic kernels below walk the window with one dependent load at a time: 0.9-1.1 TB/s).
template <bool FP16>
__global__ void __launch_bounds__(256) box3_kernel(DV in, DV out, int accumulate, uint32_t total, unsigned long long magic_c8) {
  const uint32_t c8n = (uint32_t)(in.C >> 3);
  const long long row_pitch = (long long)(in.W + 2 * in.pw) * in.ld;
  for (uint32_t t = blockIdx.x * 256u + threadIdx.x; t < total; t += gridDim.x * 256u) {
    const uint32_t m = fast_div(t, magic_c8);
    const int c8 = (int)(t - m * c8n);
    const uint32_t t2 = fast_div(m, in.magic_w);
    const int w = (int)(m - t2 * (uint32_t)in.W);
    const uint32_t nn = fast_div(t2, in.magic_h);
    const int h = (int)(t2 - nn * (uint32_t)in.H), n = (int)nn;
    const uint16_t* base = in.p + pix_off(in, n, h - 1, w - 1) + c8 * 8;      // dereferenced only where the tap is inside the image
    uint4 v[9];
    bool ok[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        ok[r * 3 + s2] = (unsigned)(h - 1 + r) < (unsigned)in.H && (unsigned)(w - 1 + s2) < (unsigned)in.W;
        if (ok[r * 3 + s2]) v[r * 3 + s2] = ld16(base + r * row_pitch + s2 * in.ld);
      }
    float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      if (ok[tap]) {
        const uint32_t u[4] = {v[tap].x, v[tap].y, v[tap].z, v[tap].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_act2(u[j], FP16 ? 1 : 0);
          g[2 * j] += f.x;
          g[2 * j + 1] += f.y;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= (1.f / 9.f);
    uint16_t* dst = out.p + pix_off(out, n, h, w) + c8 * 8;
    if (accumulate) {
      float o[8];
      load8(dst, FP16 ? 1 : 0, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += o[j];
    }
    store8(dst, FP16 ? 1 : 0, g);
  }
}

// plain average pool (count_include_pad=True, divisor k*k): F.avg_pool2d in Inception blocks / aux head
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(DV x, DV y, int k, int stride, int pad, int P, int Q, uint32_t total,
                                                          unsigned long long magic_c8, int fp16) {
  const uint32_t c8n = (uint32_t)(x.C >> 3);
  const float inv = 1.f / (float)(k * k);
  for (uint32_t t = blockIdx.x * 256u + threadIdx.x; t < total; t += gridDim.x * 256u) {
    const uint32_t m = fast_div(t, magic_c8);
    const int c8 = (int)(t - m * c8n);
    const uint32_t t2 = fast_div(m, y.magic_w);
    const int oq = (int)(m - t2 * (uint32_t)Q);
    const uint32_t nn = fast_div(t2, y.magic_h);
    const int op = (int)(t2 - nn * (uint32_t)P), n = (int)nn;
    const int h0 = op * stride - pad, w0 = oq * stride - pad;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < k; ++r) {
      const int hh = h0 + r;
      if (hh < 0 || hh >= x.H) continue;
      for (int s = 0; s < k; ++s) {
        const int ww = w0 + s;
        if (ww < 0 || ww >= x.W) continue;
        float v[8];
        load8(x.p + pix_off(x, n, hh, ww) + c8 * 8, fp16, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += v[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= inv;
    store8(y.p + pix_off(y, n, op, oq) + c8 * 8, fp16, a);
  }
}

// out[n, p*sh, q*sw, :] = in[n, p, q, :]  (out is zero elsewhere and stays so): the data gradient
// of a strided conv is the stride-1 transposed conv of the zero-dilated output gradient.
__global__ void __launch_bounds__(256) dilate_kernel(DV in, DV out, int sh, int sw, long long total) {
  const int c8n = in.C >> 3;
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
    const long long m = t / c8n;
    const int c8 = (int)(t - m * c8n);
    const int q = (int)(m % in.W);
    const long long t2 = m / in.W;
    const int p = (int)(t2 % in.H), n = (int)(t2 / in.H);
    const uint4 v = *reinterpret_cast<const uint4*>(in.p + pix_off(in, n, p, q) + c8 * 8);
    *reinterpret_cast<uint4*>(out.p + pix_off(out, n, p * sh, q * sw) + c8 * 8) = v;
  }
}

// float32 NCHW [B, Cin, H, W] -> 16-bit NHWC view with C >= Cin channels (extra channels zero):
// the stem's input in the layout the tensor-core weight-gradient kernel reads.
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, int Cin, DV out, long long total, int fp16) {
  for (long long m = (long long)blockIdx.x * 256 + threadIdx.x; m < total; m += (long long)gridDim.x * 256) {
    const int w = (int)(m % out.W);
    const long long t2 = m / out.W;
    const int h = (int)(t2 % out.H), n = (int)(t2 / out.H);
    const long long plane = (long long)out.H * out.W;
    const float* src = in + (long long)n * Cin * plane + (long long)h * out.W + w;
    uint16_t* dst = out.p + pix_off(out, n, h, w);
    for (int c0 = 0; c0 < out.C; c0 += 8) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (c0 + j) < Cin ? __ldg(src + (long long)(c0 + j) * plane) : 0.f;
      store8(dst + c0, fp16, f);
    }
  }
}

// Patch matrix of the first convolution: out[n, p, q, (r*kw + s)*3 + c] = scale[c] * in[n, c, p*st - pad + r, q*st - pad + s] + shift[c]
// (0 outside the image and for k >= kh*kw*3).  With it the Cin = 3 stem is a K = kh*kw*3 GEMM for the
// tensor-core forward and weight-gradient kernels (a 1x1 convolution over the patch tensor).
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ in, int H, int W, DV out, int kh, int kw, int stride,
                                                          int pad, float s0, float s1, float s2, float b0, float b1, float b2,
                                                          uint32_t total, unsigned long long magic_k8, int fp16) {
  const uint32_t k8n = (uint32_t)(out.C >> 3);
  const int kmax = kh * kw * 3;
  const long long plane = (long long)H * W;
  for (uint32_t idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const uint32_t m = fast_div(idx, magic_k8);
    const int k0 = (int)(idx - m * k8n) * 8;
    const uint32_t t = fast_div(m, out.magic_w);
    const int q = (int)(m - t * (uint32_t)out.W);
    const uint32_t n = fast_div(t, out.magic_h);
    const int p = (int)(t - n * (uint32_t)out.H);
    const float* src = in + (long long)n * 3 * plane;
    float f[8];
    // (tap, channel) of k0 by division once, then incrementally
    int tap = k0 / 3, c = k0 - tap * 3;
    int r = tap / kw, sx = tap - r * kw;
    const int h0 = p * stride - pad, w0 = q * stride - pad;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = 0.f;
      if (k0 + j < kmax) {
        const int hh = h0 + r, ww = w0 + sx;
        if ((unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W) {
          const float x = __ldg(src + c * plane + (long long)hh * W + ww);
          v = c == 0 ? fmaf(x, s0, b0) : c == 1 ? fmaf(x, s1, b1) : fmaf(x, s2, b2);
        }
      }
      f[j] = v;
      if (++c == 3) {
        c = 0;
        if (++sx == kw) { sx = 0; ++r; }
      }
    }
    store8(out.p + pix_off(out, (int)n, p, q) + k0, fp16, f);
  }
}

// The same patch matrix, tiled through shared memory: a block owns kTileQ output pixels of one output row.  It loads the
// input patch those pixels need (3 channels x kh rows x ((kTileQ-1)*stride + kw) columns) with row-contiguous reads, applies the
// affine and the 16-bit conversion ONCE per input element, and then assembles the [pixel][k] rows from shared memory through a
// k -> patch-offset table, writing 16-byte pieces that are contiguous across the whole tile (the gather version above issues 8
// scattered global loads per 16 output bytes: 1.1 ms for ResNet-50's 976 MB patch matrix at batch 256).
constexpr int kTileQ = 64;
__global__ void __launch_bounds__(256) stem_im2col_tiled_kernel(const float* __restrict__ in, int H, int W, DV out, int kh, int kw, int stride,
                                                                int pad, float s0, float s1, float s2, float b0, float b1, float b2,
                                                                unsigned long long magic_k8, int fp16) {
  extern __shared__ uint16_t sm16[];
  const int span = (kTileQ - 1) * stride + kw;            // input columns a tile touches
  uint16_t* patch = sm16;                                  // [3][kh][span] 16-bit activations (0 outside the image)
  uint16_t* lut = sm16 + ((3 * kh * span + 7) & ~7);       // [out.C] k -> offset in patch (0xFFFF: zero column)
  const int n = blockIdx.z, p = blockIdx.y, q0 = blockIdx.x * kTileQ;
  const int tid = threadIdx.x;
  const int kmax = kh * kw * 3;
  const long long plane = (long long)H * W;
  const float* src = in + (long long)n * 3 * plane;
  const int h0 = p * stride - pad, w0 = q0 * stride - pad;
  // one warp per (channel, filter row): lanes walk the input columns (row-contiguous reads, no per-element divisions)
  for (int cr = tid >> 5; cr < 3 * kh; cr += 8) {
    const int c = cr / kh, r = cr - c * kh;
    const int hh = h0 + r;
    const bool row_ok = (unsigned)hh < (unsigned)H;
    const float sc = c == 0 ? s0 : c == 1 ? s1 : s2, sh = c == 0 ? b0 : c == 1 ? b1 : b2;
    const float* rowp = src + c * plane + (long long)hh * W;
    for (int x = tid & 31; x < span; x += 32) {
      const int ww = w0 + x;
      float v = 0.f;
      if (row_ok && (unsigned)ww < (unsigned)W) v = fmaf(__ldg(rowp + ww), sc, sh);
      patch[cr * span + x] = (uint16_t)(pack_act2(v, 0.f, fp16) & 0xFFFFu);
    }
  }
  for (int k = tid; k < out.C; k += 256) {
    uint16_t o = 0xFFFFu;
    if (k < kmax) {
      const int tap = k / 3, c = k - tap * 3;
      const int r = tap / kw, sx = tap - r * kw;
      o = (uint16_t)((c * kh + r) * span + sx);
    }
    lut[k] = o;
  }
  __syncthreads();
  const uint32_t k8n = (uint32_t)(out.C >> 3);
  const int nq = min(kTileQ, out.W - q0);
  const uint32_t items = (uint32_t)nq * k8n;
  uint16_t* orow = out.p + pix_off(out, n, p, q0);
  for (uint32_t it = tid; it < items; it += 256) {
    const uint32_t ql = fast_div(it, magic_k8);
    const int k0 = (int)(it - ql * k8n) * 8;
    const int xo = (int)ql * stride;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t l0 = lut[k0 + 2 * j], l1 = lut[k0 + 2 * j + 1];
      const uint32_t e0 = l0 == 0xFFFFu ? 0u : patch[l0 + xo];
      const uint32_t e1 = l1 == 0xFFFFu ? 0u : patch[l1 + xo];
      w[j] = e0 | (e1 << 16);
    }
    *reinterpret_cast<uint4*>(orow + (long long)ql * out.ld + k0) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ------------------------------------------------------------------------------------------
// Classifier head, train mode: global average pool (x dropout scale) -> Linear -> CE loss.
// ------------------------------------------------------------------------------------------
// One CTA per image.  pooled[b, c] = dropscale[b, c] * mean_hw x[b, hw, c]; logits = W pooled + bias;
// loss_b = logsumexp(logits) - logits[label]; dlogits = weight/B * (softmax - onehot);
// loss_acc += weight/B * loss_b.
__global__ void __launch_bounds__(256) head_train_fwd_kernel(DV x, const float* __restrict__ dropscale, const float* __restrict__ W,
                                                             const float* __restrict__ bias, const long long* __restrict__ labels,
                                                             int n_classes, int batch, float loss_weight, float* __restrict__ pooled_out,
                                                             float* __restrict__ logits_out, float* __restrict__ dlogits,
                                                             float* __restrict__ loss_acc, int fp16) {
  extern __shared__ float sm[];
  float* pooled = sm;            // [C]
  float* logits = sm + x.C;      // [n_classes]
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HW = x.H * x.W;
  const float inv = 1.f / (float)HW;
  for (int c8 = tid; c8 < (x.C >> 3); c8 += 256) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int px = 0; px < HW; ++px) {
      float v[8];
      load8(x.p + pix_off(x, img, px / x.W, px % x.W) + c8 * 8, fp16, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
      const float ds = dropscale ? dropscale[(long long)img * x.C + c] : 1.f;
      const float pv = a[j] * inv * ds;
      pooled[c] = pv;
      pooled_out[(long long)img * x.C + c] = pv;
    }
  }
  __syncthreads();
  for (int k = warp; k < n_classes; k += 8) {
    const float4* wrow = reinterpret_cast<const float4*>(W + (long long)k * x.C);
    float s = 0.f;
    for (int c4 = lane; c4 < (x.C >> 2); c4 += 32) {
      const float4 w = __ldg(wrow + c4);
      const float4 xv = *reinterpret_cast<const float4*>(pooled + c4 * 4);
      s = fmaf(w.x, xv.x, s);
      s = fmaf(w.y, xv.y, s);
      s = fmaf(w.z, xv.z, s);
      s = fmaf(w.w, xv.w, s);
    }
    s = warp_sum(s);
    if (lane == 0) logits[k] = s + bias[k];
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int k = lane; k < n_classes; k += 32) mx = fmaxf(mx, logits[k]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < n_classes; k += 32) sum += expf(logits[k] - mx);
    sum = warp_sum(sum);
    const int label = (int)labels[img];
    const float wb = loss_weight / (float)batch;
    for (int k = lane; k < n_classes; k += 32) {
      const float l = logits[k];
      const float pr = expf(l - mx) / sum;
      if (logits_out) logits_out[(long long)img * n_classes + k] = l;
      dlogits[(long long)img * n_classes + k] = wb * (pr - (k == label ? 1.f : 0.f));
    }
    if (lane == 0 && label >= 0 && label < n_classes) atomicAdd(loss_acc, wb * (logf(sum) + mx - logits[label]));
  }
}

// dW[k, c] += sum_b dlogits[b, k] * pooled[b, c];  db[k] += sum_b dlogits[b, k]
__global__ void __launch_bounds__(256) head_wgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ pooled, int batch,
                                                         int n_classes, int C, float* __restrict__ dW, float* __restrict__ db) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t < (long long)n_classes * C) {
    const int k = (int)(t / C), c = (int)(t - (long long)k * C);
    float s = 0.f;
    for (int b = 0; b < batch; ++b) s = fmaf(__ldg(dlogits + (long long)b * n_classes + k), __ldg(pooled + (long long)b * C + c), s);
    dW[t] += s;
  }
  if (t < n_classes) {
    float s = 0.f;
    for (int b = 0; b < batch; ++b) s += dlogits[(long long)b * n_classes + t];
    db[t] += s;
  }
}

// dx[b, hw, c] (+)= dropscale[b, c] / HW * sum_k dlogits[b, k] * W[k, c]
__global__ void __launch_bounds__(256) head_dgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ W,
                                                         const float* __restrict__ dropscale, int n_classes, DV dx, int accumulate,
                                                         long long total, int fp16) {
  const int c8n = dx.C >> 3;
  const int HW = dx.H * dx.W;
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
    const int b = (int)(t / c8n), c8 = (int)(t - (long long)b * c8n);
    float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < n_classes; ++k) {
      const float dl = __ldg(dlogits + (long long)b * n_classes + k);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + (long long)k * dx.C + c8 * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + (long long)k * dx.C + c8 * 8 + 4));
      g[0] = fmaf(dl, w0.x, g[0]); g[1] = fmaf(dl, w0.y, g[1]); g[2] = fmaf(dl, w0.z, g[2]); g[3] = fmaf(dl, w0.w, g[3]);
      g[4] = fmaf(dl, w1.x, g[4]); g[5] = fmaf(dl, w1.y, g[5]); g[6] = fmaf(dl, w1.z, g[6]); g[7] = fmaf(dl, w1.w, g[7]);
    }
    const float inv = 1.f / (float)HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= inv * (dropscale ? dropscale[(long long)b * dx.C + c8 * 8 + j] : 1.f);
    for (int px = 0; px < HW; ++px) {
      uint16_t* dst = dx.p + pix_off(dx, b, px / dx.W, px % dx.W) + c8 * 8;
      float o[8];
      if (accumulate) {
        load8(dst, fp16, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += g[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = g[j];
      }
      store8(dst, fp16, o);
    }
  }
}

// inverted-dropout scale per element: 0 with probability p, else 1/(1-p).  Counter-based
// (splitmix64 of seed + index) -- its own stream; torch's Philox sequence is not reproduced.
__global__ void __launch_bounds__(256) dropout_scale_kernel(float* __restrict__ out, long long n, float p, unsigned long long seed) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    const float u = (float)(x >> 40) * (1.0f / 16777216.0f);
    out[i] = u < p ? 0.f : 1.f / (1.f - p);
  }
}

// ------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults: no weight decay, no amsgrad) over a flat fp32 arena.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                   float bc1, float bc2_sqrt, float grad_scale) {
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 w4 = reinterpret_cast<float4*>(w)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float* wp = &w4.x; float* mp = &m4.x; float* vp = &v4.x; const float* gp = &g4.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gp[j] * grad_scale;
      mp[j] = b1 * mp[j] + (1.f - b1) * gj;
      vp[j] = b2 * vp[j] + (1.f - b2) * gj * gj;
      const float denom = sqrtf(vp[j]) / bc2_sqrt + eps;
      wp[j] -= (lr / bc1) * (mp[j] / denom);
    }
    reinterpret_cast<float4*>(w)[i] = w4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
  }
  // tail (n not a multiple of 4)
  const long long i = n4 * 4 + (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) {
    const float gj = g[i] * grad_scale;
    m[i] = b1 * m[i] + (1.f - b1) * gj;
    v[i] = b2 * v[i] + (1.f - b2) * gj * gj;
    w[i] -= (lr / bc1) * (m[i] / (sqrtf(v[i]) / bc2_sqrt + eps));
  }
}

// fp32 master weights [Cout, taps, Cin] -> (1) forward operand [Cout_pad, taps*Cin_pad] 16-bit,
// (2) data-gradient operand [Cin_padN, taps*Cout_padK] 16-bit with the taps reversed:
//     wd[ci, (taps-1-t)*Cout_padK + co] = w[co, t, ci]   (conv of dz with the flipped, transposed filter)
__global__ void __launch_bounds__(256) conv_repack_kernel(const float* __restrict__ w, int Cout, int taps, int Cin, uint16_t* __restrict__ wf,
                                                          int cin_pad, uint16_t* __restrict__ wd, int cout_padk, int fp16) {
  const long long total = (long long)Cout * taps * Cin;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int ci = (int)(i % Cin);
    const long long r = i / Cin;
    const int t = (int)(r % taps), co = (int)(r / taps);
    const float v = w[i];
    uint16_t h;
    if (fp16) {
      const __half hv = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
      h = *reinterpret_cast<const uint16_t*>(&hv);
    } else {
      const __nv_bfloat16 bv = __float2bfloat16_rn(v);
      h = *reinterpret_cast<const uint16_t*>(&bv);
    }
    if (wf) wf[(long long)co * taps * cin_pad + (long long)t * cin_pad + ci] = h;
    if (wd) wd[(long long)ci * taps * cout_padk + (long long)(taps - 1 - t) * cout_padk + co] = h;
  }
}

// every conv of a network in ONE launch: blockIdx.y = conv, blockIdx.x strides over its 32 (co) x 32 (ci) tiles per filter tap.
// A tile is read once, coalesced along ci (the master layout [co][tap][ci]); the forward operand [co][tap][ci_pad] is written
// straight away (same order), the data-gradient operand [ci][taps-1-tap][co_pad] goes through a shared-memory transpose so that
// its stores are coalesced along co.  No per-element divisions: one 32-bit tile decode per 1024 elements.
// (The first version walked the elements linearly with three 64-bit divisions each and scattered the transposed store:
// 0.43-0.50 ms per step for 24 M weights; this one moves the same 190 MB at streaming rate.)
__global__ void __launch_bounds__(256) conv_repack_batch_kernel(const ifcb_repack_item* __restrict__ items, int fp16) {
  __shared__ uint16_t tile[32][33];
  const ifcb_repack_item it = items[blockIdx.y];
  const float* __restrict__ w = it.d_master;
  uint16_t* __restrict__ wf = reinterpret_cast<uint16_t*>(it.d_wfwd);
  uint16_t* __restrict__ wd = reinterpret_cast<uint16_t*>(it.d_wdgrad);
  const int Cin = it.Cin, Cout = it.Cout, taps = it.taps;
  const int cit_n = (Cin + 31) >> 5, cot_n = (Cout + 31) >> 5;
  const int tiles = taps * cot_n * cit_n;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
    const int cit = tl % cit_n, r = tl / cit_n;
    const int cot = r % cot_n, t = r / cot_n;
    const int ci = cit * 32 + tx;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int co = cot * 32 + ty + 8 * k;
      uint16_t h = 0;
      if (co < Cout && ci < Cin) {
        const float v = w[((long long)co * taps + t) * Cin + ci];
        if (fp16) {
          const __half hv = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
          h = *reinterpret_cast<const uint16_t*>(&hv);
        } else {
          const __nv_bfloat16 bv = __float2bfloat16_rn(v);
          h = *reinterpret_cast<const uint16_t*>(&bv);
        }
        if (wf) wf[((long long)co * taps + t) * it.Cin_pad + ci] = h;
      }
      tile[ty + 8 * k][tx] = h;
    }
    __syncthreads();
    if (wd) {
      const int co = cot * 32 + tx;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ci2 = cit * 32 + ty + 8 * k;
        if (co < Cout && ci2 < Cin) wd[((long long)ci2 * taps + (taps - 1 - t)) * it.Cout_padk + co] = tile[tx][ty + 8 * k];
      }
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace ifcb

using namespace ifcb;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define DT_OK(dt) ((dt) == IFCB_ACT_BF16 || (dt) == IFCB_ACT_FP16)

extern "C" int ifcb_memset_zero(void* d, int64_t bytes, void* stream) {
  IFCB_ARG_CHECK(d != nullptr && bytes >= 0, "memset_zero: bad argument");
  IFCB_CUDA_CHECK(cudaMemsetAsync(d, 0, (size_t)bytes, STREAM(stream)));
  return 0;
}

static int reduce_rows(int C) {
  const int c8n = C / 8;
  int rows = 256 / c8n;
  return rows < 1 ? 1 : rows;
}

// grid of the streaming BN kernels: blocks of `rows` pixel rows, each thread kUnroll rows per sweep;
// at most 8 resident blocks per SM (2048 threads), at least one sweep of work per block
static int stream_grid(long long M, int rows) {
  long long g = (M + (long long)rows * kUnroll - 1) / ((long long)rows * kUnroll);
  const long long cap = (long long)sm_count() * 8;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

// Reduction kernels: channel blocks of <= 64 channels (c8b groups of 8) x pixel splits.  A block streams at least ~32 KB per
// tensor before its 2 * 8 * c8b float64 atomics; the whole grid is at most 8 blocks per SM.
struct ReduceCfg {
  int c8b, rows;
  dim3 grid;
};
static ReduceCfg reduce_cfg(long long M, int C) {
  ReduceCfg r;
  static const bool whole = getenv("IFCB_BN_REDUCE_WHOLE") != nullptr;      // A/B switch: every block takes all channels
  const int c8n = C / 8;
  if (whole || c8n <= 16) {
    r.c8b = c8n;                                         // narrow tensors: a block reads whole pixel rows
  } else {
    r.c8b = 8;                                           // 128 / 64 / 32 / 16-byte channel blocks: always whole sectors
    while (c8n % r.c8b) r.c8b >>= 1;
  }
  r.rows = 256 / r.c8b;
  if (r.rows < 1) r.rows = 1;
  const int ychunks = (C / 8) / r.c8b;
  long long xs = (M + (long long)r.rows * kUnroll - 1) / ((long long)r.rows * kUnroll);          // one sweep per block at most
  const long long cap = ((long long)sm_count() * 8 + ychunks - 1) / ychunks;
  if (xs > cap) xs = cap;
  const long long blk_bytes = whole ? (256 << 10) : (32 << 10);
  const long long by_bytes = (M * 16 * r.c8b + blk_bytes - 1) / blk_bytes;
  if (xs > by_bytes) xs = by_bytes;
  if (xs < 1) xs = 1;
  r.grid = dim3((unsigned)xs, (unsigned)ychunks, 1);
  return r;
}

extern "C" int ifcb_bn_stats(const ifcb_view* z, int batch, int dtype, float eps, float momentum, double* d_acc, float* d_mean,
                             float* d_invstd, float* d_running_mean, float* d_running_var, void* stream) {
  IFCB_ARG_CHECK(view_ok(z) && batch > 0 && DT_OK(dtype), "bn_stats: bad view / batch / dtype");
  IFCB_ARG_CHECK(z->C <= 2048, "bn_stats: C=%d > 2048", z->C);
  IFCB_ARG_CHECK(d_acc && d_mean && d_invstd, "bn_stats: null pointer");
  const long long M = (long long)batch * z->H * z->W;
  const ReduceCfg rc = reduce_cfg(M, z->C);
  DV zz = dv(z);
  IFCB_ARG_CHECK(M < (1ll << 31) / (z->C / 8), "bn_stats: tensor too large for 32-bit indexing");
  FinalizeArgs fin{};
  fin.count = (double)M;
  fin.eps = eps;
  fin.momentum = momentum;
  fin.mean_out = d_mean;
  fin.invstd_out = d_invstd;
  fin.run_mean = d_running_mean;
  fin.run_var = d_running_var;
  fin.counter = reinterpret_cast<unsigned int*>(d_acc + 7680);       // one arrival counter per channel block (<= 256), behind the coefficients
  if (det_enabled()) {
    const long long need = 4ll * rc.grid.x * rc.grid.y * 16 * rc.c8b;
    fin.det_part = static_cast<float*>(det_workspace(need));
    IFCB_ARG_CHECK(fin.det_part != nullptr, "bn_stats: the deterministic workspace is smaller than %lld bytes", need);
  }
  channel_reduce_kernel<0, kMaskNone><<<rc.grid, 256, 0, STREAM(stream)>>>(zz, zz, zz, nullptr, nullptr, nullptr, nullptr, M, rc.rows, dtype, d_acc,
                                                                           fin, rc.c8b);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_bn_apply(const ifcb_view* z, const ifcb_view* out, const ifcb_view* residual, int batch, int dtype,
                             const float* d_mean, const float* d_invstd, const float* d_gamma, const float* d_beta, int relu,
                             void* stream) {
  IFCB_ARG_CHECK(view_ok(z) && view_ok(out) && batch > 0 && DT_OK(dtype), "bn_apply: bad view / batch / dtype");
  IFCB_ARG_CHECK(out->C == z->C && out->H == z->H && out->W == z->W, "bn_apply: output extent differs");
  IFCB_ARG_CHECK(!residual || (view_ok(residual) && residual->C == z->C && residual->H == z->H && residual->W == z->W),
                 "bn_apply: residual extent differs");
  IFCB_ARG_CHECK(d_mean && d_invstd && d_gamma && d_beta, "bn_apply: null pointer");
  IFCB_ARG_CHECK(z->C <= 2048, "bn_apply: C=%d > 2048", z->C);
  const long long M = (long long)batch * z->H * z->W;
  IFCB_ARG_CHECK(M < (1ll << 31), "bn_apply: tensor too large for 32-bit indexing");
  const int rows = reduce_rows(z->C);
  DV zz = dv(z);
  if (residual)
    bn_apply_kernel<true><<<stream_grid(M, rows), 256, 0, STREAM(stream)>>>(zz, dv(out), dv(residual), d_mean, d_invstd, d_gamma, d_beta, relu,
                                                                             M, rows, dtype);
  else
    bn_apply_kernel<false><<<stream_grid(M, rows), 256, 0, STREAM(stream)>>>(zz, dv(out), zz, d_mean, d_invstd, d_gamma, d_beta, relu, M, rows,
                                                                              dtype);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_bn_apply_sums(const ifcb_view* z, const ifcb_view* out, const ifcb_view* residual, int batch, int dtype,
                                  const double* d_sums, float eps, float momentum, float* d_mean, float* d_invstd,
                                  float* d_running_mean, float* d_running_var, const float* d_gamma, const float* d_beta, int relu,
                                  void* stream) {
  IFCB_ARG_CHECK(view_ok(z) && view_ok(out) && batch > 0 && DT_OK(dtype), "bn_apply_sums: bad view / batch / dtype");
  IFCB_ARG_CHECK(out->C == z->C && out->H == z->H && out->W == z->W, "bn_apply_sums: output extent differs");
  IFCB_ARG_CHECK(!residual || (view_ok(residual) && residual->C == z->C && residual->H == z->H && residual->W == z->W),
                 "bn_apply_sums: residual extent differs");
  IFCB_ARG_CHECK(d_sums && d_mean && d_invstd && d_gamma && d_beta, "bn_apply_sums: null pointer");
  IFCB_ARG_CHECK(z->C <= 2048, "bn_apply_sums: C=%d > 2048", z->C);
  const long long M = (long long)batch * z->H * z->W;
  IFCB_ARG_CHECK(M < (1ll << 31), "bn_apply_sums: tensor too large for 32-bit indexing");
  const int rows = reduce_rows(z->C);
  DV zz = dv(z);
  if (residual)
    bn_apply_sums_kernel<true><<<stream_grid(M, rows), 256, 0, STREAM(stream)>>>(zz, dv(out), dv(residual), d_sums, (double)M, eps, momentum,
                                                                                  d_mean, d_invstd, d_running_mean, d_running_var, d_gamma,
                                                                                  d_beta, relu, M, rows, dtype);
  else
    bn_apply_sums_kernel<false><<<stream_grid(M, rows), 256, 0, STREAM(stream)>>>(zz, dv(out), zz, d_sums, (double)M, eps, momentum, d_mean,
                                                                                   d_invstd, d_running_mean, d_running_var, d_gamma, d_beta,
                                                                                   relu, M, rows, dtype);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static int bn_backward_impl(const ifcb_view* dy, const ifcb_view* a, const ifcb_view* z, const ifcb_view* dz,
                            const ifcb_view* dres, int dres_accumulate, int relu, int batch, int dtype, const float* d_mean,
                            const float* d_invstd, const float* d_gamma, const float* d_beta, double* d_acc, float* d_dgamma,
                            float* d_dbeta, int dz_accumulate, void* stream);

extern "C" int ifcb_bn_backward(const ifcb_view* dy, const ifcb_view* a, const ifcb_view* z, const ifcb_view* dz,
                                const ifcb_view* dres, int dres_accumulate, int relu, int batch, int dtype, const float* d_mean,
                                const float* d_invstd, const float* d_gamma, const float* d_beta, double* d_acc, float* d_dgamma,
                                float* d_dbeta, void* stream) {
  return bn_backward_impl(dy, a, z, dz, dres, dres_accumulate, relu, batch, dtype, d_mean, d_invstd, d_gamma, d_beta, d_acc, d_dgamma,
                          d_dbeta, 0, stream);
}

extern "C" int ifcb_bn_backward_accumulate(const ifcb_view* dy, const ifcb_view* z, const ifcb_view* dz, int relu, int batch, int dtype,
                                           const float* d_mean, const float* d_invstd, const float* d_gamma, const float* d_beta,
                                           double* d_acc, float* d_dgamma, float* d_dbeta, void* stream) {
  return bn_backward_impl(dy, nullptr, z, dz, nullptr, 0, relu, batch, dtype, d_mean, d_invstd, d_gamma, d_beta, d_acc, d_dgamma, d_dbeta, 1,
                          stream);
}

static int bn_backward_impl(const ifcb_view* dy, const ifcb_view* a, const ifcb_view* z, const ifcb_view* dz,
                            const ifcb_view* dres, int dres_accumulate, int relu, int batch, int dtype, const float* d_mean,
                            const float* d_invstd, const float* d_gamma, const float* d_beta, double* d_acc, float* d_dgamma,
                            float* d_dbeta, int dz_accumulate, void* stream) {
  IFCB_ARG_CHECK(view_ok(dy) && view_ok(z) && view_ok(dz) && batch > 0 && DT_OK(dtype), "bn_backward: bad view / batch / dtype");
  IFCB_ARG_CHECK(z->C <= 2048, "bn_backward: C=%d > 2048", z->C);
  IFCB_ARG_CHECK(dy->C == z->C && dz->C == z->C && dy->H == z->H && dy->W == z->W && dz->H == z->H && dz->W == z->W,
                 "bn_backward: extents differ");
  IFCB_ARG_CHECK(!a || (view_ok(a) && a->C == z->C && a->H == z->H && a->W == z->W), "bn_backward: mask extent differs");
  IFCB_ARG_CHECK(!dres || (view_ok(dres) && dres->C == z->C && dres->H == z->H && dres->W == z->W), "bn_backward: dres extent differs");
  IFCB_ARG_CHECK(d_mean && d_invstd && d_gamma && d_beta && d_acc, "bn_backward: null pointer");
  IFCB_ARG_CHECK(!relu || !dres || a, "bn_backward: ReLU after a residual add needs the forward output `a` for the mask");
  const long long M = (long long)batch * z->H * z->W;
  const long long total = M * (z->C / 8);
  IFCB_ARG_CHECK(total < (1ll << 31), "bn_backward: tensor too large for 32-bit indexing");
  // ReLU mask: recomputed from z (bit-identical to the forward, saves reading `a`) unless a residual was added
  const int mask_mode = !relu ? kMaskNone : (dres ? kMaskFromA : kMaskFromZ);
  const int rows = reduce_rows(z->C);
  DV zz = dv(z), dyy = dv(dy);
  // coefficient floats live behind the accumulator area of the LARGEST layer (2 * 2048 float64), never inside it:
  // a smaller layer's coefficients must not land where a wider layer accumulates next
  float* coef = reinterpret_cast<float*>(d_acc + 2 * 2048);
  const int grid = stream_grid(M, rows);
  cudaStream_t st = STREAM(stream);
  DV av = a ? dv(a) : zz, dzv = dv(dz), drv = dres ? dv(dres) : zz;
  const int res_mode = dres ? (dres_accumulate ? 2 : 1) : 0;
  FinalizeArgs fin{};
  fin.count = (double)M;
  fin.coef = coef;
  fin.dgamma = d_dgamma;
  fin.dbeta = d_dbeta;
  fin.counter = reinterpret_cast<unsigned int*>(d_acc + 7680);
  const ReduceCfg rc = reduce_cfg(M, z->C);
  if (det_enabled()) {
    const long long need = 4ll * rc.grid.x * rc.grid.y * 16 * rc.c8b;
    fin.det_part = static_cast<float*>(det_workspace(need));
    IFCB_ARG_CHECK(fin.det_part != nullptr, "bn_backward: the deterministic workspace is smaller than %lld bytes", need);
  }
#define IFCB_REDUCE1(MK) channel_reduce_kernel<1, MK><<<rc.grid, 256, 0, st>>>(zz, dyy, av, d_mean, d_invstd, d_gamma, d_beta, M, rc.rows, dtype, d_acc, fin, rc.c8b)
  if (mask_mode == kMaskFromZ) IFCB_REDUCE1(kMaskFromZ);
  else if (mask_mode == kMaskFromA) IFCB_REDUCE1(kMaskFromA);
  else IFCB_REDUCE1(kMaskNone);
#undef IFCB_REDUCE1
#define IFCB_BWD_APPLY(MK, RS) \
  bn_bwd_apply_kernel<MK, RS><<<grid, 256, 0, st>>>(dyy, av, zz, dzv, drv, d_mean, d_invstd, d_gamma, d_beta, coef, M, rows, dtype)
  if (dz_accumulate) {
    IFCB_ARG_CHECK(!dres && dz->d != dy->d, "bn_backward_accumulate: no residual route, dz must not alias dy");
    if (mask_mode == kMaskFromZ) bn_bwd_apply_kernel<kMaskFromZ, 0, true><<<grid, 256, 0, st>>>(dyy, av, zz, dzv, drv, d_mean, d_invstd, d_gamma, d_beta, coef, M, rows, dtype);
    else bn_bwd_apply_kernel<kMaskNone, 0, true><<<grid, 256, 0, st>>>(dyy, av, zz, dzv, drv, d_mean, d_invstd, d_gamma, d_beta, coef, M, rows, dtype);
  } else
  if (mask_mode == kMaskFromZ) IFCB_BWD_APPLY(kMaskFromZ, 0);               // (a residual always comes with kMaskFromA / kMaskNone)
  else if (mask_mode == kMaskFromA) { if (res_mode == 2) IFCB_BWD_APPLY(kMaskFromA, 2); else if (res_mode == 1) IFCB_BWD_APPLY(kMaskFromA, 1); else IFCB_BWD_APPLY(kMaskFromA, 0); }
  else { if (res_mode == 2) IFCB_BWD_APPLY(kMaskNone, 2); else if (res_mode == 1) IFCB_BWD_APPLY(kMaskNone, 1); else IFCB_BWD_APPLY(kMaskNone, 0); }
#undef IFCB_BWD_APPLY
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static int pool_out(int in, int k, int s, int p) { return (in + 2 * p - k) / s + 1; }
// torch's ceil_mode output size: one more window when it still starts inside the (left-padded) input
static int pool_out_ceil(int in, int k, int s, int p) {
  int o = (in + 2 * p - k + s - 1) / s + 1;
  if ((o - 1) * s >= in + p) --o;
  return o;
}
static bool pool_dim_ok(int out, int in, int k, int s, int p) { return out == pool_out(in, k, s, p) || out == pool_out_ceil(in, k, s, p); }

extern "C" int ifcb_bias_relu_backward(const ifcb_view* dy, const ifcb_view* a, const ifcb_view* dz, int relu, int batch, int dtype,
                                       double* d_acc, float* d_dbias, void* stream) {
  IFCB_ARG_CHECK(view_ok(dy) && view_ok(dz) && batch > 0 && DT_OK(dtype), "bias_relu_backward: bad view / batch / dtype");
  IFCB_ARG_CHECK(dz->C == dy->C && dz->H == dy->H && dz->W == dy->W, "bias_relu_backward: extents differ");
  IFCB_ARG_CHECK(!relu || (view_ok(a) && a->C == dy->C && a->H == dy->H && a->W == dy->W), "bias_relu_backward: ReLU needs the forward output");
  IFCB_ARG_CHECK(d_acc && d_dbias && dy->C <= 4096, "bias_relu_backward: null pointer / C > 4096");
  const long long M = (long long)batch * dy->H * dy->W;
  IFCB_ARG_CHECK(M * (dy->C / 8) < (1ll << 31), "bias_relu_backward: tensor too large for 32-bit indexing");
  const ReduceCfg rc = reduce_cfg(M, dy->C);
  float* det = nullptr;
  if (det_enabled()) {
    const long long need = 4ll * rc.grid.x * rc.grid.y * 8 * rc.c8b;
    det = static_cast<float*>(det_workspace(need));
    IFCB_ARG_CHECK(det != nullptr, "bias_relu_backward: the deterministic workspace is smaller than %lld bytes", need);
  }
  unsigned int* counter = reinterpret_cast<unsigned int*>(d_acc + 7680);
  DV dyy = dv(dy), dzz = dv(dz), aa = relu ? dv(a) : dyy;
  if (relu) bias_relu_bwd_kernel<true><<<rc.grid, 256, 0, STREAM(stream)>>>(dyy, aa, dzz, M, rc.rows, dtype, d_acc, d_dbias, counter, det, rc.c8b);
  else bias_relu_bwd_kernel<false><<<rc.grid, 256, 0, STREAM(stream)>>>(dyy, aa, dzz, M, rc.rows, dtype, d_acc, d_dbias, counter, det, rc.c8b);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_scale_elems(const ifcb_view* x, const ifcb_view* y, const float* d_scale, int batch, int dtype, void* stream) {
  IFCB_ARG_CHECK(view_ok(x) && view_ok(y) && d_scale && batch > 0 && DT_OK(dtype), "scale_elems: bad argument");
  IFCB_ARG_CHECK(x->C == y->C && x->H == y->H && x->W == y->W, "scale_elems: extents differ");
  IFCB_ARG_CHECK((reinterpret_cast<uintptr_t>(d_scale) & 15) == 0, "scale_elems: scale must be 16-byte aligned");
  const long long total = (long long)batch * x->H * x->W * (x->C / 8);
  IFCB_ARG_CHECK(total < (1ll << 31), "scale_elems: tensor too large for 32-bit indexing");
  scale_elems_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(x), dv(y), d_scale, (uint32_t)total, div_magic(x->C / 8), dtype);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_maxpool_fwd_train(const ifcb_view* x, const ifcb_view* y, uint8_t* d_idx, int batch, int k, int stride, int pad,
                                      int dtype, void* stream) {
  IFCB_ARG_CHECK(view_ok(x) && view_ok(y) && d_idx && batch > 0 && DT_OK(dtype), "maxpool_fwd_train: bad argument");
  IFCB_ARG_CHECK(k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && pad < k, "maxpool_fwd_train: bad window");
  const int P = y->H, Q = y->W;        // floor or ceil_mode extent (torch.nn.MaxPool2d(ceil_mode=True): squeezenet.py)
  IFCB_ARG_CHECK(y->C == x->C && pool_dim_ok(P, x->H, k, stride, pad) && pool_dim_ok(Q, x->W, k, stride, pad),
                 "maxpool_fwd_train: output extent %dx%dx%d does not fit input %dx%dx%d", y->H, y->W, y->C, x->H, x->W, x->C);
  const long long total = (long long)batch * P * Q * (x->C / 8);
  IFCB_ARG_CHECK(total < (1ll << 31), "maxpool_fwd_train: tensor too large for 32-bit indexing");
  if (k == 3 && dtype)
    maxpool3_fwd_kernel<true><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(x), dv(y), d_idx, stride, pad, P, Q, (uint32_t)total,
                                                                                 div_magic(x->C / 8));
  else if (k == 3)
    maxpool3_fwd_kernel<false><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(x), dv(y), d_idx, stride, pad, P, Q, (uint32_t)total,
                                                                                  div_magic(x->C / 8));
  else
    maxpool_fwd_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(x), dv(y), d_idx, k, stride, pad, P, Q, (uint32_t)total,
                                                                          div_magic(x->C / 8), dtype);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_maxpool_bwd(const ifcb_view* dy, const uint8_t* d_idx, const ifcb_view* dx, int accumulate, int batch, int k,
                                int stride, int pad, int dtype, void* stream) {
  IFCB_ARG_CHECK(view_ok(dy) && view_ok(dx) && d_idx && batch > 0 && DT_OK(dtype), "maxpool_bwd: bad argument");
  IFCB_ARG_CHECK(k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && pad < k, "maxpool_bwd: bad window");
  const int P = dy->H, Q = dy->W;
  IFCB_ARG_CHECK(dy->C == dx->C && pool_dim_ok(P, dx->H, k, stride, pad) && pool_dim_ok(Q, dx->W, k, stride, pad), "maxpool_bwd: gradient extent differs");
  const long long total = (long long)batch * dx->H * dx->W * (dx->C / 8);
  IFCB_ARG_CHECK(total < (1ll << 31), "maxpool_bwd: tensor too large for 32-bit indexing");
  if (k == 3 && stride == 2 && dtype)
    maxpool3s2_bwd_kernel<true><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(dy), d_idx, dv(dx), accumulate, pad, P, Q,
                                                                                   (uint32_t)total, div_magic(dx->C / 8));
  else if (k == 3 && stride == 2)
    maxpool3s2_bwd_kernel<false><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(dy), d_idx, dv(dx), accumulate, pad, P, Q,
                                                                                    (uint32_t)total, div_magic(dx->C / 8));
  else
    pool_bwd_kernel<false><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(dy), d_idx, dv(dx), accumulate, k, stride, pad, P, Q,
                                                                              (uint32_t)total, div_magic(dx->C / 8), dtype);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_avgpool_fwd(const ifcb_view* x, const ifcb_view* y, int batch, int k, int stride, int pad, int dtype, void* stream) {
  IFCB_ARG_CHECK(view_ok(x) && view_ok(y) && batch > 0 && DT_OK(dtype), "avgpool_fwd: bad argument");
  IFCB_ARG_CHECK(k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && pad < k, "avgpool_fwd: bad window");
  const int P = pool_out(x->H, k, stride, pad), Q = pool_out(x->W, k, stride, pad);
  IFCB_ARG_CHECK(y->C == x->C && y->H == P && y->W == Q, "avgpool_fwd: output extent differs");
  const long long total = (long long)batch * P * Q * (x->C / 8);
  IFCB_ARG_CHECK(total < (1ll << 31), "avgpool_fwd: tensor too large for 32-bit indexing");
  if (k == 3 && stride == 1 && pad == 1) {
    if (dtype) box3_kernel<true><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(x), dv(y), 0, (uint32_t)total, div_magic(x->C / 8));
    else box3_kernel<false><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(x), dv(y), 0, (uint32_t)total, div_magic(x->C / 8));
  } else {
    avgpool_fwd_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(x), dv(y), k, stride, pad, P, Q, (uint32_t)total,
                                                                          div_magic(x->C / 8), dtype);
  }
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_avgpool_bwd(const ifcb_view* dy, const ifcb_view* dx, int accumulate, int batch, int k, int stride, int pad,
                                int dtype, void* stream) {
  IFCB_ARG_CHECK(view_ok(dy) && view_ok(dx) && batch > 0 && DT_OK(dtype), "avgpool_bwd: bad argument");
  IFCB_ARG_CHECK(k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && pad < k, "avgpool_bwd: bad window");
  const int P = pool_out(dx->H, k, stride, pad), Q = pool_out(dx->W, k, stride, pad);
  IFCB_ARG_CHECK(dy->C == dx->C && dy->H == P && dy->W == Q, "avgpool_bwd: gradient extent differs");
  const long long total = (long long)batch * dx->H * dx->W * (dx->C / 8);
  IFCB_ARG_CHECK(total < (1ll << 31), "avgpool_bwd: tensor too large for 32-bit indexing");
  if (k == 3 && stride == 1 && pad == 1) {        // the gradient of a 3x3/s1/p1 box mean is the same box mean of dy
    if (dtype) box3_kernel<true><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(dy), dv(dx), accumulate, (uint32_t)total, div_magic(dx->C / 8));
    else box3_kernel<false><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(dy), dv(dx), accumulate, (uint32_t)total, div_magic(dx->C / 8));
  } else {
    pool_bwd_kernel<true><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(dy), nullptr, dv(dx), accumulate, k, stride, pad, P, Q,
                                                                             (uint32_t)total, div_magic(dx->C / 8), dtype);
  }
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_dilate(const ifcb_view* in, const ifcb_view* out, int batch, int stride_h, int stride_w, void* stream) {
  IFCB_ARG_CHECK(view_ok(in) && view_ok(out) && batch > 0, "dilate: bad argument");
  IFCB_ARG_CHECK(stride_h >= 1 && stride_w >= 1 && out->C == in->C && out->H >= (in->H - 1) * stride_h + 1 &&
                 out->W >= (in->W - 1) * stride_w + 1, "dilate: output too small");
  const long long total = (long long)batch * in->H * in->W * (in->C / 8);
  dilate_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(dv(in), dv(out), stride_h, stride_w, total);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_nchw_to_nhwc(const float* d_in, int Cin, const ifcb_view* out, int batch, int dtype, void* stream) {
  IFCB_ARG_CHECK(d_in && view_ok(out) && batch > 0 && Cin > 0 && Cin <= out->C && DT_OK(dtype), "nchw_to_nhwc: bad argument");
  const long long total = (long long)batch * out->H * out->W;
  nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(d_in, Cin, dv(out), total, dtype);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_stem_im2col(const float* d_in, int H, int W, const ifcb_view* out, int batch, int kh, int kw, int stride, int pad,
                                const float* h_scale, const float* h_shift, int dtype, void* stream) {
  IFCB_ARG_CHECK(d_in && view_ok(out) && batch > 0 && DT_OK(dtype), "stem_im2col: bad argument");
  IFCB_ARG_CHECK(kh >= 1 && kw >= 1 && stride >= 1 && pad >= 0 && out->C >= kh * kw * 3, "stem_im2col: bad window / channel count");
  IFCB_ARG_CHECK(out->H == (H + 2 * pad - kh) / stride + 1 && out->W == (W + 2 * pad - kw) / stride + 1, "stem_im2col: output extent differs");
  const long long total = (long long)batch * out->H * out->W * (out->C / 8);
  IFCB_ARG_CHECK(total < (1ll << 31), "stem_im2col: tensor too large for 32-bit indexing");
  const float sc[3] = {h_scale ? h_scale[0] : 1.f, h_scale ? h_scale[1] : 1.f, h_scale ? h_scale[2] : 1.f};
  const float sh[3] = {h_shift ? h_shift[0] : 0.f, h_shift ? h_shift[1] : 0.f, h_shift ? h_shift[2] : 0.f};
  const int span = (kTileQ - 1) * stride + kw;
  const int smem = (((3 * kh * span + 7) & ~7) + out->C) * 2;
  static const bool gather = getenv("IFCB_STEM_IM2COL_GATHER") != nullptr;       // A/B switch: the first (gather) kernel
  if (!gather && smem <= 48 * 1024 && out->pad_w == 0 && batch <= 65535 && out->H <= 65535) {
    dim3 grid((out->W + kTileQ - 1) / kTileQ, out->H, batch);
    stem_im2col_tiled_kernel<<<grid, 256, smem, STREAM(stream)>>>(d_in, H, W, dv(out), kh, kw, stride, pad, sc[0], sc[1], sc[2], sh[0], sh[1],
                                                                   sh[2], div_magic(out->C / 8), dtype);
  } else {
    stem_im2col_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(d_in, H, W, dv(out), kh, kw, stride, pad, sc[0], sc[1], sc[2], sh[0],
                                                                          sh[1], sh[2], (uint32_t)total, div_magic(out->C / 8), dtype);
  }
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_head_train_fwd(const ifcb_view* x, int batch, int dtype, const float* d_dropscale, const float* d_weight,
                                   const float* d_bias, const int64_t* d_labels, int n_classes, float loss_weight, float* d_pooled,
                                   float* d_logits, float* d_dlogits, float* d_loss, void* stream) {
  IFCB_ARG_CHECK(view_ok(x) && batch > 0 && DT_OK(dtype), "head_train_fwd: bad view / batch / dtype");
  IFCB_ARG_CHECK(d_weight && d_bias && d_labels && d_pooled && d_dlogits && d_loss, "head_train_fwd: null pointer");
  IFCB_ARG_CHECK(n_classes > 0 && (x->C + n_classes) * 4 <= 200 * 1024, "head_train_fwd: n_classes out of range");
  IFCB_ARG_CHECK((reinterpret_cast<uintptr_t>(d_weight) & 15) == 0, "head_train_fwd: weight must be 16-byte aligned");
  const int smem = (x->C + n_classes) * 4;
  if (smem > 48 * 1024) {   // per-device function attribute
    static bool done[64] = {};
    int dev = 0;
    IFCB_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !done[dev]) {
      IFCB_CUDA_CHECK(cudaFuncSetAttribute(head_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      if (dev >= 0 && dev < 64) done[dev] = true;
    }
  }
  head_train_fwd_kernel<<<batch, 256, smem, STREAM(stream)>>>(dv(x), d_dropscale, d_weight, d_bias, reinterpret_cast<const long long*>(d_labels),
                                                               n_classes, batch, loss_weight, d_pooled, d_logits, d_dlogits, d_loss, dtype);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_head_bwd(const ifcb_view* dx, int dx_accumulate, int batch, int dtype, const float* d_dropscale, const float* d_weight,
                             const float* d_pooled, const float* d_dlogits, int n_classes, float* d_dweight, float* d_dbias, void* stream) {
  IFCB_ARG_CHECK(view_ok(dx) && batch > 0 && DT_OK(dtype), "head_bwd: bad view / batch / dtype");
  IFCB_ARG_CHECK(d_weight && d_pooled && d_dlogits && d_dweight && d_dbias && n_classes > 0, "head_bwd: null pointer");
  const long long nw = (long long)n_classes * dx->C;
  head_wgrad_kernel<<<(int)((nw + 255) / 256), 256, 0, STREAM(stream)>>>(d_dlogits, d_pooled, batch, n_classes, dx->C, d_dweight, d_dbias);
  const long long total = (long long)batch * (dx->C / 8);
  head_dgrad_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(d_dlogits, d_weight, d_dropscale, n_classes, dv(dx), dx_accumulate,
                                                                       total, dtype);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_dropout_scale(float* d_out, int64_t n, float p, uint64_t seed, void* stream) {
  IFCB_ARG_CHECK(d_out && n > 0 && p >= 0.f && p < 1.f, "dropout_scale: bad argument");
  dropout_scale_kernel<<<grid_for(n, 256), 256, 0, STREAM(stream)>>>(d_out, n, p, seed);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_adam_step(float* d_param, const float* d_grad, float* d_m, float* d_v, int64_t n, float lr, float beta1,
                              float beta2, float eps, int step, float grad_scale, void* stream) {
  IFCB_ARG_CHECK(d_param && d_grad && d_m && d_v && n > 0 && step >= 1, "adam_step: bad argument");
  IFCB_ARG_CHECK(((reinterpret_cast<uintptr_t>(d_param) | reinterpret_cast<uintptr_t>(d_grad) | reinterpret_cast<uintptr_t>(d_m) |
                   reinterpret_cast<uintptr_t>(d_v)) & 15) == 0, "adam_step: arenas must be 16-byte aligned");
  const float bc1 = 1.f - (float)pow((double)beta1, (double)step);
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  adam_kernel<<<grid_for((n + 3) / 4 + 4, 256), 256, 0, STREAM(stream)>>>(d_param, d_grad, d_m, d_v, n, lr, beta1, beta2, eps, bc1,
                                                                           bc2_sqrt, grad_scale);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_conv_repack(const float* d_master, int Cout, int taps, int Cin, void* d_wfwd, int Cin_pad, void* d_wdgrad,
                                int Cout_padk, int dtype, void* stream) {
  IFCB_ARG_CHECK(d_master && Cout > 0 && taps > 0 && Cin > 0 && DT_OK(dtype), "conv_repack: bad argument");
  IFCB_ARG_CHECK(!d_wfwd || Cin_pad >= Cin, "conv_repack: Cin_pad < Cin");
  IFCB_ARG_CHECK(!d_wdgrad || Cout_padk >= Cout, "conv_repack: Cout_padk < Cout");
  const long long total = (long long)Cout * taps * Cin;
  conv_repack_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(d_master, Cout, taps, Cin, reinterpret_cast<uint16_t*>(d_wfwd), Cin_pad,
                                                                        reinterpret_cast<uint16_t*>(d_wdgrad), Cout_padk, dtype);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int ifcb_conv_repack_batch(const ifcb_repack_item* d_items, int n_items, int dtype, void* stream) {
  IFCB_ARG_CHECK(d_items && n_items > 0 && n_items <= 65535 && DT_OK(dtype), "conv_repack_batch: bad argument");
  conv_repack_batch_kernel<<<dim3(sm_count(), n_items), 256, 0, STREAM(stream)>>>(d_items, dtype);   // the largest convs (2.4 M weights) need the whole GPU; blocks of small ones exit at once
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

