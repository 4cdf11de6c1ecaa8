// K2 -- Conv2d + folded BatchNorm (+residual) + ReLU as an implicit GEMM on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a only.
//
// Replaces torchvision BasicConv2d (inception.py:398-407) / ResNet conv-bn-relu as
// reached from NeustonModel.forward (reference neuston_models.py:66-68).
//
// GEMM view: D[M, N] = A[M, K] * B[N, K]^T, N = Cout, K = kh*kw*Cin, B = packed weights
// (tiled TMA).  A is never materialised in HBM; two operand-feeding schemes:
//
//  * IM2COL (any stride): M = batch*P*Q output pixels.  One TMA *im2col* load per (filter
//    tap, 64-channel block) fetches the 128 x 64 operand tile of 128 consecutive output
//    pixels (padding = TMA zero fill, stride = TMA traversal stride).
//
//  * WINDOW (stride 1): the input lives in HBM physically zero-padded, [N, Hp, Wp, C], and
//    is viewed as a plain 2-D matrix [N*Hp*Wp, C].  GEMM row i is anchored at padded pixel
//    i (top-left of the receptive field), so filter tap (r, s) needs rows i + r*Wp + s: a
//    CONSTANT shift.  Per 64-channel block ONE tiled TMA load brings rows
//    [i0, i0 + 128*m + halo) into shared memory and every tap is just a UMMA descriptor
//    whose start address is shifted by (r*Wp + s) rows -- the patch is read from L2 once
//    instead of kh*kw times, and m (1..4) accumulators of 128 rows share each weight tile.
//    Rows whose anchor falls in the padding produce junk that the epilogue drops.
//
// Persistent, warp-specialised CTA (1 per SM, 320 threads):
//   warp 0     TMA producer (one elected lane)
//   warp 1     TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2-9  epilogue: tcgen05.ld -> scale/shift (+residual) -> ReLU -> 16-bit ->
//              swizzled smem tile -> coalesced 16-byte stores (two warps per TMEM lane
//              quadrant, alternating 64-column groups)
// Pipelines: smem rings (full/empty mbarriers) TMA <-> MMA; two TMEM accumulator buffers
// (256 columns each) MMA <-> epilogue.
#include "layers.cuh"
#include "ptx.cuh"

namespace ifcb {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                         // 16-bit elements = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;       // 320
constexpr int kTmemCols = 512;
constexpr int kAccBufCols = 256;
constexpr int kBarBytes = 512;

// barriers + scale/shift (2 x cout_pad floats) + one 4 KB staging tile per epilogue warp
__host__ __device__ constexpr int epilogue_smem(int cout_pad) {
  return kBarBytes + 8 * cout_pad + kEpiWarps * 4096;
}

struct RowMap {           // where a GEMM row lands
  int n, p, q;
  bool valid;
};

__device__ __forceinline__ RowMap map_row(const ConvKernelParams& p, long long i) {
  RowMap r;
  const int n = (int)(i / p.rows_per_img);
  const int rem = (int)(i - (long long)n * p.rows_per_img);
  r.n = n;
  r.p = rem / p.row_w;
  r.q = rem - r.p * p.row_w;
  r.valid = (i < p.rows) && (r.p < p.P) && (r.q < p.Q);
  return r;
}

// ---------------------------------------------------------------------------------------
// Epilogue for one 128-row accumulator (TMEM columns [tcol, tcol + tile_n)), executed by
// the two warps of one lane quadrant (half = 0/1 takes the even/odd 64-column groups).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_128rows(const ConvKernelParams& p, uint32_t taddr, long long row0,
                                                 int n_lo, int lane, int half, uint32_t stage_addr,
                                                 const float* s_scale, const float* s_shift) {
  const RowMap rm = map_row(p, row0 + lane);
  long long res_row = 0;
  if (p.residual != nullptr)
    res_row = ((long long)rm.n * (p.P + 2 * p.res_pad_h) + rm.p + p.res_pad_h) * (p.Q + 2 * p.res_pad_w) + rm.q + p.res_pad_w;
  const int n_hi = n_lo + p.tile_n;
  int gcount = 0;
  for (int si = 0; si < p.n_seg; ++si) {
    const int g_lo = max(n_lo, p.seg_begin[si]), g_hi = min(n_hi, p.seg_end[si]);
    if (g_lo >= g_hi) continue;
    const int relu = p.seg_relu[si];
    // destination row of this lane's GEMM row inside the (possibly padded) output tensor
    const int Hd = p.P + 2 * p.seg_pad_h[si], Wd = p.Q + 2 * p.seg_pad_w[si];
    const int drow = rm.valid ? ((rm.n * Hd + rm.p + p.seg_pad_h[si]) * Wd + rm.q + p.seg_pad_w[si]) : -1;
    const int ld = p.seg_ld[si];
    for (int g0 = g_lo; g0 < g_hi; g0 += 64, ++gcount) {
      if ((gcount & 1) != half) continue;
      const int nch = min(4, (g_hi - g0) >> 4);
      uint32_t v[4][16];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        if (ch < nch) ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)(g0 + ch * 16 - n_lo), v[ch]);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if (ch < nch) {
          const int n = g0 + ch * 16;
          float y[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + n + 4 * j);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + n + 4 * j);
            y[4 * j + 0] = fmaf(__uint_as_float(v[ch][4 * j + 0]), sc.x, sh.x);
            y[4 * j + 1] = fmaf(__uint_as_float(v[ch][4 * j + 1]), sc.y, sh.y);
            y[4 * j + 2] = fmaf(__uint_as_float(v[ch][4 * j + 2]), sc.z, sh.z);
            y[4 * j + 3] = fmaf(__uint_as_float(v[ch][4 * j + 3]), sc.w, sh.w);
          }
          if (p.residual != nullptr && rm.valid) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + res_row * p.res_ld + n);
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
      }
      __syncwarp();
      // coalesced write-out: ppr 16-byte pieces per row (2, 4, 6 or 8), 32 rows
      const int ppr = 2 * nch;
      __nv_bfloat16* gout = p.seg_out[si] + (g0 - p.seg_begin[si]);
      for (int idx = lane; idx < 32 * ppr; idx += 32) {
        int r;
        if (ppr == 8) r = idx >> 3;
        else if (ppr == 4) r = idx >> 2;
        else if (ppr == 2) r = idx >> 1;
        else r = (idx * 171) >> 10;                      // idx / 6 for idx < 192
        const int pc = idx - r * ppr;
        const int dr = __shfl_sync(0xffffffffu, drow, r);
        const uint4 val = ptx::ld_shared_v4(stage_addr + (uint32_t)r * 128u + ((((uint32_t)pc) ^ (uint32_t)(r & 7)) << 4));
        if (dr >= 0) *reinterpret_cast<uint4*>(gout + (long long)dr * ld + pc * 8) = val;
      }
      __syncwarp();
    }
  }
}

template <bool WINDOW>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B operand tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = p.tile_n * kBlockK * 2;
  // IM2COL: `stages` x [A 16 KB | B]          WINDOW: a_slots x A patch, then `stages` x B
  uint8_t* a_base = smem;
  uint8_t* b_base = WINDOW ? smem + (size_t)p.a_slots * p.a_slot_bytes : smem + kATileBytes;
  const int b_stride = WINDOW ? b_tile_bytes : kATileBytes + b_tile_bytes;
  uint8_t* bar_base = WINDOW ? b_base + (size_t)p.stages * b_tile_bytes : smem + (size_t)p.stages * b_stride;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);     // [stages]  (B, or A+B for IM2COL)
  uint64_t* empty_bar = full_bar + 12;                            // [stages]
  uint64_t* a_full = full_bar + 24;                               // [a_slots] (WINDOW)
  uint64_t* a_empty = full_bar + 28;
  uint64_t* tmem_full = full_bar + 32;                            // [2]
  uint64_t* tmem_empty = full_bar + 34;                           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 36);
  float* s_scale = reinterpret_cast<float*>(bar_base + kBarBytes);
  float* s_shift = s_scale + p.cout_pad;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_shift + p.cout_pad);      // 8 x 4 KB, 128-byte aligned
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
    for (int s = 0; s < p.a_slots; ++s) {
      ptx::mbar_init(a_full + s, 1);
      ptx::mbar_init(a_empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tmem_full + s, 1);
      ptx::mbar_init(tmem_empty + s, kEpiWarps);      // one arrive per epilogue warp
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

  const int tile_rows = kBlockM * p.m_sub;
  const int m_tiles = (int)((p.rows + tile_rows - 1) / tile_rows);
  const int total_tiles = m_tiles * p.n_tiles;
  const int taps = p.kh * p.kw;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, aslot = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        if (WINDOW) {
          const int i0 = m_tile * tile_rows;
          for (int cb = 0; cb < p.cblocks; ++cb) {
            ptx::mbar_wait(a_empty + aslot, aphase ^ 1);
            uint8_t* dst = a_base + (size_t)aslot * p.a_slot_bytes;
            if (p.debug_flags & 1) {
              ptx::mbar_arrive(a_full + aslot);
            } else {
              ptx::mbar_arrive_expect_tx(a_full + aslot, (uint32_t)(p.n_boxes * p.box_rows * 128));
              for (int b = 0; b < p.n_boxes; ++b)
                ptx::tma_load_2d(dst + (size_t)b * p.box_rows * 128, &tmap_a, a_full + aslot, cb * kBlockK,
                                 i0 + b * p.box_rows);
            }
            if (++aslot == p.a_slots) { aslot = 0; aphase ^= 1; }
            for (int t = 0; t < taps; ++t) {
              ptx::mbar_wait(empty_bar + stage, phase ^ 1);
              if (p.debug_flags & 1) {
                ptx::mbar_arrive(full_bar + stage);
              } else {
                ptx::mbar_arrive_expect_tx(full_bar + stage, (uint32_t)b_tile_bytes);
                ptx::tma_load_2d(b_base + (size_t)stage * b_stride, &tmap_b, full_bar + stage,
                                 (t * p.cblocks + cb) * kBlockK, n_tile * p.tile_n);
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        } else {
          const int m0 = m_tile * kBlockM;
          const int img = m0 / p.rows_per_img;
          const int rem = m0 - img * p.rows_per_img;
          const int op = rem / p.row_w, oq = rem - op * p.row_w;
          const int w0 = oq * p.stride_w - p.pad_w;
          const int h0 = op * p.stride_h - p.pad_h;
          for (int r = 0; r < p.kh; ++r) {
            for (int s = 0; s < p.kw; ++s) {
              for (int cb = 0; cb < p.cblocks; ++cb) {
                ptx::mbar_wait(empty_bar + stage, phase ^ 1);
                uint8_t* a_dst = a_base + (size_t)stage * b_stride;
                if (p.debug_flags & 1) {
                  ptx::mbar_arrive(full_bar + stage);
                } else {
                  ptx::mbar_arrive_expect_tx(full_bar + stage, (uint32_t)(kATileBytes + b_tile_bytes));
                  ptx::tma_load_im2col_4d(a_dst, &tmap_a, full_bar + stage, cb * kBlockK, w0, h0, img,
                                          (uint16_t)s, (uint16_t)r);
                  ptx::tma_load_2d(a_dst + kATileBytes, &tmap_b, full_bar + stage,
                                   ((r * p.kw + s) * p.cblocks + cb) * kBlockK, n_tile * p.tile_n);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_f16(kBlockM, p.tile_n, p.fp16);
      int stage = 0, aslot = 0;
      uint32_t phase = 0, aphase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        ptx::mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccBufCols);
        if (WINDOW) {
          for (int cb = 0; cb < p.cblocks; ++cb) {
            ptx::mbar_wait(a_full + aslot, aphase);
            ptx::tc_fence_after();
            const uint32_t a_addr = ptx::smem_u32(a_base + (size_t)aslot * p.a_slot_bytes);
            const int ksteps = (cb == p.cblocks - 1) ? p.last_ksteps : (kBlockK / 16);
            for (int t = 0; t < taps; ++t) {
              const int r = t / p.kw, s = t - r * p.kw;
              ptx::mbar_wait(full_bar + stage, phase);
              ptx::tc_fence_after();
              const uint64_t bdesc = ptx::umma_desc_k_sw128(ptx::smem_u32(b_base + (size_t)stage * b_stride));
              const uint32_t shift_rows = (uint32_t)(p.win_shift0 + r * p.row_w + s);
              for (int j = 0; j < p.m_sub; ++j) {
                const uint32_t a_start = a_addr + (shift_rows + (uint32_t)(j * kBlockM)) * 128u;
                const uint64_t adesc = ptx::umma_desc_k_sw128_shifted(a_start, p.desc_base_offset_mode);
                if (!(p.debug_flags & 2))
                  for (int k = 0; k < ksteps; ++k)
                    ptx::umma_f16(d_tmem + (uint32_t)(j * p.tile_n), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k),
                                  idesc, (cb > 0 || t > 0 || k > 0) ? 1u : 0u);
              }
              ptx::umma_commit(empty_bar + stage);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            ptx::umma_commit(a_empty + aslot);
            if (++aslot == p.a_slots) { aslot = 0; aphase ^= 1; }
          }
        } else {
          const int kblocks = taps * p.cblocks;
          for (int kb = 0; kb < kblocks; ++kb) {
            ptx::mbar_wait(full_bar + stage, phase);
            ptx::tc_fence_after();
            const uint32_t a_addr = ptx::smem_u32(a_base + (size_t)stage * b_stride);
            const uint64_t adesc = ptx::umma_desc_k_sw128(a_addr);
            const uint64_t bdesc = ptx::umma_desc_k_sw128(a_addr + kATileBytes);
            const int cb = kb % p.cblocks;
            const int ksteps = (cb == p.cblocks - 1) ? p.last_ksteps : (kBlockK / 16);
            if (!(p.debug_flags & 2))
              for (int k = 0; k < ksteps; ++k)
                ptx::umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                              (kb > 0 || k > 0) ? 1u : 0u);
            ptx::umma_commit(empty_bar + stage);     // frees the smem slot when these MMAs retire
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        ptx::umma_commit(tmem_full + acc);           // accumulators ready for the epilogue
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;                // which of the quadrant's two warps
    const uint32_t stage_addr = ptx::smem_u32(s_stage + (warp - 2) * 4096);
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      ptx::mbar_wait(tmem_full + acc, acc_phase);
      ptx::tc_fence_after();
      for (int j = 0; j < p.m_sub && !(p.debug_flags & 4); ++j) {
        const long long row0 = (long long)m_tile * tile_rows + j * kBlockM + quad * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kAccBufCols + j * p.tile_n);
        epilogue_128rows(p, taddr, row0, n_tile * p.tile_n, lane, half, stage_addr, s_scale, s_shift);
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

constexpr int kSmemBudget = 227 * 1024;

}  // namespace

// Shared-memory plan of a layer.  Returns false if nothing fits.
bool conv_plan_smem(ConvKernelParams& kp, bool window, int halo_rows) {
  const int b_tile = kp.tile_n * kBlockK * 2;
  const int fixed = epilogue_smem(kp.cout_pad) + 1024;
  if (!window) {
    kp.m_sub = 1;
    kp.a_slots = 0;
    kp.a_slot_bytes = 0;
    kp.box_rows = kp.n_boxes = 0;
    int s = (kSmemBudget - fixed) / (kATileBytes + b_tile);
    if (s > 10) s = 10;
    kp.stages = s;
    return s >= 2;
  }
  // WINDOW: choose m (accumulators per tile) as large as TMEM double buffering and smem allow
  for (int m = 4; m >= 1; m >>= 1) {
    if (m > kp.m_sub_cap) continue;
    if (m * kp.tile_n > kAccBufCols) continue;
    const int rows = kBlockM * m + halo_rows;
    const int n_boxes = (rows + 255) / 256;
    int box_rows = ((rows + n_boxes - 1) / n_boxes + 7) & ~7;
    const int slot = (n_boxes * box_rows * 128 + 1023) & ~1023;
    for (int slots = kp.a_slots_pref; slots >= 1; --slots) {
      const int left = kSmemBudget - fixed - slots * slot;
      int s = left / b_tile;
      if (s > 10) s = 10;
      const int need = slots >= 2 ? 3 : 4;
      if (s >= need) {
        kp.m_sub = m;
        kp.a_slots = slots;
        kp.a_slot_bytes = slot;
        kp.box_rows = box_rows;
        kp.n_boxes = n_boxes;
        kp.stages = s;
        return true;
      }
      if (m > 1) break;      // prefer a smaller m with two slots over a single-slot large m
    }
  }
  return false;
}

int conv_smem_bytes(const ConvKernelParams& kp, bool window) {
  const int b_tile = kp.tile_n * kBlockK * 2;
  const int ops = window ? kp.a_slots * kp.a_slot_bytes + kp.stages * b_tile : kp.stages * (kATileBytes + b_tile);
  return ops + epilogue_smem(kp.cout_pad) + 1024;
}

int launch_conv(const ConvLayer& L, int batch, cudaStream_t stream) {
  ConvKernelParams p = L.kp;
  p.rows = (long long)batch * p.rows_per_img;
  const int tile_rows = kBlockM * p.m_sub;
  const long long m_tiles = (p.rows + tile_rows - 1) / tile_rows;
  const long long total = m_tiles * p.n_tiles;
  if (total == 0) return 0;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  const int smem = conv_smem_bytes(p, L.window);
  static int attr_smem[2] = {0, 0};
  if (smem > attr_smem[L.window ? 1 : 0]) {
    if (L.window)
      IFCB_CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else
      IFCB_CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem[L.window ? 1 : 0] = smem;
  }
  if (L.window)
    conv_umma_kernel<true><<<grid, kThreads, smem, stream>>>(L.tmap_a, L.tmap_b, p);
  else
    conv_umma_kernel<false><<<grid, kThreads, smem, stream>>>(L.tmap_a, L.tmap_b, p);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// Debug probe (tests only): issue ONE im2col TMA load with a layer's tensor map
// and copy the raw 16 KB shared-memory tile (still 128B-swizzled) to global, so
// the TMA im2col semantics can be checked independently of the MMA path.
// ---------------------------------------------------------------------------------
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
