// K2 -- Conv2d + folded BatchNorm + ReLU as an implicit GEMM on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a only.
//
// Replaces torchvision BasicConv2d (inception.py:398-407) / ResNet conv-bn-relu as
// reached from NeustonModel.forward (reference neuston_models.py:66-68).
//
// GEMM view: D[M, N] = A[M, K] * B[N, K]^T with
//   M = batch * P * Q output pixels (NHWC order), N = Cout, K = kh*kw*Cin.
//   A is never materialised: one TMA *im2col* load per (filter tap, 64-channel
//   block) fetches the 128 x 64 bf16 operand tile for 128 consecutive output
//   pixels straight from the NHWC activation tensor (padding = TMA zero fill,
//   stride = TMA traversal stride) into 128B-swizzled shared memory.
//   B (packed weights) comes in with a plain tiled TMA load.
//
// Persistent, warp-specialised CTA (1 per SM, 192 threads):
//   warp 0    TMA producer      (one elected lane)
//   warp 1    TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2-5 epilogue: tcgen05.ld -> scale/shift (+residual) -> ReLU -> bf16
//             -> 16-byte stores into the (possibly concatenated) NHWC output
// Pipelines: smem ring (full/empty mbarriers) between TMA and MMA; two TMEM
// accumulator stages (256 columns each) between MMA and epilogue, so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include "layers.cuh"
#include "ptx.cuh"

namespace ifcb {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;
constexpr int kAccStageCols = 256;
// barriers (256 B) + scale/shift (2 x cout_pad floats) + 4 x 4 KB staging tiles
__host__ __device__ constexpr int kEpilogueSmem(int cout_pad) { return 256 + 8 * cout_pad + 4 * 4096; }

__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B operand tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = p.tile_n * kBlockK * 2;
  const int stage_bytes = kATileBytes + b_tile_bytes;
  uint8_t* bar_base = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  // epilogue scratch: folded-BN scale/shift for every GEMM column of this layer, and a
  // private 32-row x 128-byte staging tile per epilogue warp (swizzled, conflict free)
  float* s_scale = reinterpret_cast<float*>(bar_base + 256);
  float* s_shift = s_scale + p.cout_pad;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_shift + p.cout_pad);      // 4 warps x 4 KB, 16-byte aligned
  for (int i = threadIdx.x; i < p.cout_pad; i += kThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar + s, 1);
      ptx::mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tmem_full + s, 1);
      ptx::mbar_init(tmem_empty + s, 4);      // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int total_tiles = m_tiles * p.n_tiles;
  const int taps = p.kh * p.kw;
  const int kblocks = taps * p.cblocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        const int m0 = m_tile * kBlockM;
        const int img = m0 / p.PQ;
        const int rem = m0 - img * p.PQ;
        const int op = rem / p.Q, oq = rem - op * p.Q;
        const int w0 = oq * p.stride_w - p.pad_w;
        const int h0 = op * p.stride_h - p.pad_h;
        for (int r = 0; r < p.kh; ++r) {
          for (int s = 0; s < p.kw; ++s) {
            for (int cb = 0; cb < p.cblocks; ++cb) {
              ptx::mbar_wait(empty_bar + stage, phase ^ 1);
              uint8_t* a_dst = smem + (size_t)stage * stage_bytes;
              uint8_t* b_dst = a_dst + kATileBytes;
              ptx::mbar_arrive_expect_tx(full_bar + stage, (uint32_t)stage_bytes);
              ptx::tma_load_im2col_4d(a_dst, &tmap_a, full_bar + stage, cb * kBlockK, w0, h0, img,
                                      (uint16_t)s, (uint16_t)r);
              ptx::tma_load_2d(b_dst, &tmap_b, full_bar + stage,
                               ((r * p.kw + s) * p.cblocks + cb) * kBlockK, n_tile * p.tile_n);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_f16(kBlockM, p.tile_n, p.fp16);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        ptx::mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStageCols);
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(full_bar + stage, phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = ptx::umma_desc_k_sw128(a_addr);
          const uint64_t bdesc = ptx::umma_desc_k_sw128(a_addr + kATileBytes);
          const int cb = kb % p.cblocks;
          const int ksteps = (cb == p.cblocks - 1) ? p.last_ksteps : (kBlockK / 16);
          for (int k = 0; k < ksteps; ++k) {
            // advance 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
            ptx::umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar + stage);     // frees the smem slot when these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(tmem_full + acc);         // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // TMEM -> registers -> affine/ReLU -> 16-bit -> per-warp swizzled smem tile -> coalesced
    // 16-byte global stores (4 full 128-byte lines per warp instruction for a 64-column group).
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
    uint8_t* stage = s_stage + (warp - 2) * 4096;
    const uint32_t stage_addr = ptx::smem_u32(stage);
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      ptx::mbar_wait(tmem_full + acc, acc_phase);
      ptx::tc_fence_after();
      const long long m_warp = (long long)m_tile * kBlockM + quad * 32;      // first row of this warp
      const long long m = m_warp + lane;
      const bool row_ok = m < p.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kAccStageCols);
      const int n_lo = n_tile * p.tile_n, n_hi = n_lo + p.tile_n;
      for (int si = 0; si < p.n_seg; ++si) {
        const int g_lo = max(n_lo, p.seg_begin[si]), g_hi = min(n_hi, p.seg_end[si]);
        const int relu = p.seg_relu[si];
        for (int g0 = g_lo; g0 < g_hi; g0 += 64) {            // group of up to 4 chunks of 16 columns
          const int nch = min(4, (g_hi - g0) >> 4);
          for (int ch = 0; ch < nch; ++ch) {
            const int n = g0 + ch * 16;
            uint32_t v[16];
            ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)(n - n_lo), v);
            ptx::tmem_ld_wait();
            float y[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 sc = *reinterpret_cast<const float4*>(s_scale + n + 4 * j);
              const float4 sh = *reinterpret_cast<const float4*>(s_shift + n + 4 * j);
              y[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), sc.x, sh.x);
              y[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), sc.y, sh.y);
              y[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), sc.z, sh.z);
              y[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), sc.w, sh.w);
            }
            if (p.residual != nullptr && row_ok) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + m * p.res_ld + n);
              const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
              const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 f = unpack_act2(rr[j], p.fp16);
                y[2 * j] += f.x;
                y[2 * j + 1] += f.y;
              }
            }
            if (relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j], 0.f);
            }
            uint4 o0, o1;
            o0.x = pack_act2(y[0], y[1], p.fp16);
            o0.y = pack_act2(y[2], y[3], p.fp16);
            o0.z = pack_act2(y[4], y[5], p.fp16);
            o0.w = pack_act2(y[6], y[7], p.fp16);
            o1.x = pack_act2(y[8], y[9], p.fp16);
            o1.y = pack_act2(y[10], y[11], p.fp16);
            o1.z = pack_act2(y[12], y[13], p.fp16);
            o1.w = pack_act2(y[14], y[15], p.fp16);
            // row `lane`, 16-byte pieces 2*ch and 2*ch+1, XOR-swizzled by (row & 7)
            const uint32_t rbase = stage_addr + (uint32_t)lane * 128u;
            const uint32_t sw = (uint32_t)(lane & 7);
            ptx::st_shared_v4(rbase + ((((uint32_t)(2 * ch)) ^ sw) << 4), o0);
            ptx::st_shared_v4(rbase + ((((uint32_t)(2 * ch + 1)) ^ sw) << 4), o1);
          }
          __syncwarp();
          // coalesced write-out: `ppr` 16-byte pieces per row, 32 rows
          const int ppr = 2 * nch;
          __nv_bfloat16* gout = p.seg_out[si] + (g0 - p.seg_begin[si]);
          const int ld = p.seg_ld[si];
          for (int idx = lane; idx < 32 * ppr; idx += 32) {
            const int r = idx / ppr, pc = idx - r * ppr;
            const uint4 val = ptx::ld_shared_v4(stage_addr + (uint32_t)r * 128u + ((((uint32_t)pc) ^ (uint32_t)(r & 7)) << 4));
            if (m_warp + r < p.M)
              *reinterpret_cast<uint4*>(gout + (m_warp + r) * ld + pc * 8) = val;
          }
          __syncwarp();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty + acc);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

int conv_smem_bytes(int tile_n, int stages, int cout_pad) {
  return stages * (kATileBytes + tile_n * kBlockK * 2) + kEpilogueSmem(cout_pad) + 1024;
}

int conv_pick_stages(int tile_n, int cout_pad) {
  const int budget = 227 * 1024;
  int s = (budget - kEpilogueSmem(cout_pad) - 1024) / (kATileBytes + tile_n * kBlockK * 2);
  if (s > 8) s = 8;
  return s;
}

int launch_conv(const ConvLayer& L, int batch, cudaStream_t stream) {
  ConvKernelParams p = L.kp;
  p.M = batch * p.PQ;
  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int total = m_tiles * p.n_tiles;
  if (total == 0) return 0;
  int grid = total < sm_count() ? total : sm_count();
  const int smem = conv_smem_bytes(p.tile_n, p.stages, p.cout_pad);
  static int attr_smem = 0;
  if (smem > attr_smem) {
    IFCB_CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  conv_umma_kernel<<<grid, kThreads, smem, stream>>>(L.tmap_a, L.tmap_b, p);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace ifcb

// ---------------------------------------------------------------------------------
// Debug probe (tests only): issue ONE im2col TMA load with a layer's tensor map
// and copy the raw 16 KB shared-memory tile (still 128B-swizzled) to global, so
// the TMA im2col semantics can be checked independently of the MMA path.
// ---------------------------------------------------------------------------------
namespace ifcb {
namespace {
__global__ void im2col_probe_kernel(const __grid_constant__ CUtensorMap tmap_a, int c, int w, int h, int n,
                                    int off_w, int off_h, uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kATileBytes);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar, kATileBytes);
    ptx::tma_load_im2col_4d(smem, &tmap_a, bar, c, w, h, n, (uint16_t)off_w, (uint16_t)off_h);
  }
  ptx::mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < kATileBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(smem)[i];
}
}  // namespace

int launch_im2col_probe(const ConvLayer& L, int c, int w, int h, int n, int off_w, int off_h, void* d_out,
                        cudaStream_t stream) {
  const int smem = kATileBytes + 1024 + 64;
  im2col_probe_kernel<<<1, 128, smem, stream>>>(L.tmap_a, c, w, h, n, off_w, off_h,
                                                reinterpret_cast<uint8_t*>(d_out));
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
}  // namespace ifcb
