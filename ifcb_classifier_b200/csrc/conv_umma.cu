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
//    tap, channel block) fetches the 128-row operand tile of 128 consecutive output
//    pixels (padding = TMA zero fill, stride = TMA traversal stride).
//
//  * WINDOW (stride 1): the input lives in HBM physically zero-padded, [N, Hp, Wp, C], and
//    is viewed as a plain 2-D matrix [N*Hp*Wp, C].  GEMM row i is anchored at padded pixel
//    i (top-left of the receptive field), so filter tap (r, s) needs rows i + r*Wp + s: a
//    CONSTANT shift.  Per channel block ONE tiled TMA load brings rows
//    [i0, i0 + 128*m + halo) into shared memory and every tap is just a UMMA descriptor
//    whose start address is shifted by (r*Wp + s) rows -- the patch is read from L2 once
//    instead of kh*kw times, and m (1..4) accumulators of 128 rows share each weight tile.
//    Weight tiles of several taps travel as one pipeline stage (one barrier round trip per
//    `b_group` taps).  Rows whose anchor falls in the padding produce junk that the
//    epilogue drops.
//
// Operand rows are one swizzle span wide: 128 bytes (64 channels, SWIZZLE_128B) or, for
// Cin <= 32, 64 bytes (32 channels, SWIZZLE_64B) -- halves the shared-memory footprint of
// the first layers.
//
// Persistent, warp-specialised CTA (1 per SM, 320 threads):
//   warp 0     TMA producer
//   warp 1     TMEM allocator + tcgen05.mma issuer
//   warps 2-9  epilogue: tcgen05.ld -> scale/shift (+residual) -> ReLU -> 16-bit ->
//              swizzled smem tile -> coalesced 16-byte stores (two warps per TMEM lane
//              quadrant, alternating 64-column groups)
// Warps 0/1 run their loops warp-uniformly and elect one lane only around the async
// instructions, so descriptors and coordinates stay in uniform registers (a lane-0 branch
// around the whole loop costs ~25 SASS instructions per tcgen05.mma, measured).
// Pipelines: smem rings (full/empty mbarriers) TMA <-> MMA; two TMEM accumulator buffers
// (256 columns each) MMA <-> epilogue.
#include "layers.cuh"
#include "ptx.cuh"

namespace ifcb {

namespace {

constexpr int kBlockM = 128;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;       // 320
constexpr int kTmemCols = 512;
constexpr int kAccBufCols = 256;
constexpr int kBarBytes = 512;
constexpr int kMaxStages = 12;

// barriers + scale/shift (2 x cout_pad floats) + one 4 KB staging tile per epilogue warp
__host__ __device__ constexpr int epilogue_smem(int cout_pad) {
  return kBarBytes + 8 * cout_pad + kEpiWarps * 4096;
}

struct RowMap {           // where a GEMM row lands
  int n, p, q;
  bool valid;
};

// floor(n / d) for 0 <= n < 2^31 through the host-computed magic = floor((2^64 - 1) / d) + 1
// (exact while n * d < 2^64; magic == 0 encodes d == 1).  ~5 integer instructions
// instead of the ~40 of an emulated 32-bit division -- map_row runs per accumulator.
__device__ __forceinline__ uint32_t fast_div(uint32_t n, unsigned long long magic) {
  if (magic == 0ull) return n;
  const uint32_t m_lo = (uint32_t)magic, m_hi = (uint32_t)(magic >> 32);
  const unsigned long long t = (unsigned long long)n * m_hi + (((unsigned long long)n * m_lo) >> 32);
  return (uint32_t)(t >> 32);
}

__device__ __forceinline__ RowMap map_row(const ConvKernelParams& p, int i) {
  RowMap r;
  const int n = (int)fast_div((uint32_t)i, p.magic_img);
  const int rem = i - n * p.rows_per_img;
  r.n = n;
  r.p = (int)fast_div((uint32_t)rem, p.magic_w);
  r.q = rem - r.p * p.row_w;
  r.valid = (i < (int)p.rows) && (r.p < p.P) && (r.q < p.Q);
  return r;
}

// two fp32 -> packed 16-bit pair (lo in bits 0-15), optional ReLU fused into the convert;
// fp16 saturates at +-65504 (satfinite)
template <bool FP16, bool RELU>
__device__ __forceinline__ uint32_t cvt_pack(float lo, float hi) {
  uint32_t r;
  if (FP16) {
    if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  }
  return r;
}

template <bool FP16, bool RELU>
__device__ __forceinline__ void pack16(const float (&y)[16], uint4& o0, uint4& o1) {
  o0.x = cvt_pack<FP16, RELU>(y[0], y[1]);
  o0.y = cvt_pack<FP16, RELU>(y[2], y[3]);
  o0.z = cvt_pack<FP16, RELU>(y[4], y[5]);
  o0.w = cvt_pack<FP16, RELU>(y[6], y[7]);
  o1.x = cvt_pack<FP16, RELU>(y[8], y[9]);
  o1.y = cvt_pack<FP16, RELU>(y[10], y[11]);
  o1.z = cvt_pack<FP16, RELU>(y[12], y[13]);
  o1.w = cvt_pack<FP16, RELU>(y[14], y[15]);
}

// coalesced write-out of one staged group: 32 rows x PPR 16-byte pieces, fully unrolled
template <int PPR>
__device__ __forceinline__ void write_out(uint32_t stage_addr, int lane, int drow, __nv_bfloat16* gout, int ld) {
  uint4 val[PPR];
  int dr[PPR], pcs[PPR];
#pragma unroll
  for (int it = 0; it < PPR; ++it) {
    const int idx = lane + 32 * it;
    const int r = PPR == 8 ? idx >> 3 : PPR == 4 ? idx >> 2 : PPR == 2 ? idx >> 1 : (idx * 171) >> 10;   // idx / PPR
    const int pc = idx - r * PPR;
    pcs[it] = pc;
    dr[it] = __shfl_sync(0xffffffffu, drow, r);
    val[it] = ptx::ld_shared_v4(stage_addr + (uint32_t)r * 128u + ((((uint32_t)pc) ^ (uint32_t)(r & 7)) << 4));
  }
#pragma unroll
  for (int it = 0; it < PPR; ++it)
    if (dr[it] >= 0) *reinterpret_cast<uint4*>(gout + (long long)dr[it] * ld + pcs[it] * 8) = val[it];
}

// ---------------------------------------------------------------------------------------
// Epilogue for one 128-row accumulator (TMEM columns [tcol, tcol + tile_n)), executed by
// the two warps of one lane quadrant; `gcount` numbers the 64-column groups of the whole
// tile and warp `half` (0/1) takes the even/odd ones.  The row mapping (two divisions) is
// only evaluated by a warp that actually owns a group of this accumulator.
// ---------------------------------------------------------------------------------------
template <bool FP16>
__device__ __forceinline__ void epilogue_128rows(const ConvKernelParams& p, uint32_t taddr, int row0, int n_lo,
                                                 int lane, int half, int& gcount, uint32_t stage_addr,
                                                 const float* s_scale, const float* s_shift) {
  const int n_hi = n_lo + p.tile_n;
  const uint32_t rbase = stage_addr + (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)(lane & 7);
  bool mapped = false;
  RowMap rm;
  long long res_row = 0;
  for (int si = 0; si < p.n_seg; ++si) {
    const int g_lo = max(n_lo, p.seg_begin[si]), g_hi = min(n_hi, p.seg_end[si]);
    if (g_lo >= g_hi) continue;
    const int ngroups = (g_hi - g_lo + 63) >> 6;
    // does this warp own any group of the segment?  (groups alternate between the two warps)
    if (ngroups == 1 && (gcount & 1) != half) { ++gcount; continue; }
    if (!mapped) {
      mapped = true;
      if (p.identity_rows) {
        rm.n = 0; rm.p = 0; rm.q = row0 + lane;
        rm.valid = (row0 + lane) < (int)p.rows;
      } else {
        rm = map_row(p, row0 + lane);
      }
      if (p.residual != nullptr)
        res_row = p.identity_rows ? (long long)(row0 + lane)
                                  : ((long long)rm.n * (p.P + 2 * p.res_pad_h) + rm.p + p.res_pad_h) * (p.Q + 2 * p.res_pad_w) + rm.q + p.res_pad_w;
    }
    const int relu = p.seg_relu[si];
    // destination row of this lane's GEMM row inside the (possibly padded) output tensor
    int drow;
    if (p.identity_rows) {
      drow = rm.valid ? rm.q : -1;
    } else {
      const int Hd = p.P + 2 * p.seg_pad_h[si], Wd = p.Q + 2 * p.seg_pad_w[si];
      drow = rm.valid ? ((rm.n * Hd + rm.p + p.seg_pad_h[si]) * Wd + rm.q + p.seg_pad_w[si]) : -1;
    }
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
              const float2 f = unpack_act2(rr[j], FP16 ? 1 : 0);
              y[2 * j] += f.x;
              y[2 * j + 1] += f.y;
            }
          }
          uint4 o0, o1;
          if (relu) pack16<FP16, true>(y, o0, o1);
          else pack16<FP16, false>(y, o0, o1);
          // row `lane`, 16-byte pieces 2*ch and 2*ch+1, XOR-swizzled by (row & 7)
          ptx::st_shared_v4(rbase + ((((uint32_t)(2 * ch)) ^ sw) << 4), o0);
          ptx::st_shared_v4(rbase + ((((uint32_t)(2 * ch + 1)) ^ sw) << 4), o1);
        }
      }
      __syncwarp();
      // coalesced write-out: 2*nch 16-byte pieces per row, 32 rows
      __nv_bfloat16* gout = p.seg_out[si] + (g0 - p.seg_begin[si]);
      if (nch == 4) write_out<8>(stage_addr, lane, drow, gout, ld);
      else if (nch == 2) write_out<4>(stage_addr, lane, drow, gout, ld);
      else if (nch == 3) write_out<6>(stage_addr, lane, drow, gout, ld);
      else write_out<2>(stage_addr, lane, drow, gout, ld);
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------
// MMA issue, specialised on the K steps per operand row (KS) and the accumulators per tile
// (MS) so that every tcgen05.mma of a filter tap is straight-line code on the uniform
// datapath: one tap = MS x KS MMAs, operands a_lo + j*jstride + 2k / b_lo + 2k.
// ---------------------------------------------------------------------------------------
template <int KS, int MS>
__device__ __forceinline__ void issue_tap(bool leader, uint32_t a_lo, uint32_t b_lo, uint32_t jstride, uint32_t d0,
                                          uint32_t dstep, uint32_t desc_hi, uint32_t idesc, uint32_t first) {
#pragma unroll
  for (int j = 0; j < MS; ++j) {
#pragma unroll
    for (int k = 0; k < KS; ++k)
      if (leader)
        ptx::umma_f16_lohi(d0 + (uint32_t)j * dstep, a_lo + (uint32_t)j * jstride + (uint32_t)(2 * k),
                           b_lo + (uint32_t)(2 * k), desc_hi, idesc, k > 0 ? 1u : first);
  }
}

// all taps [t0, t0 + nt) of one weight pipeline stage (WINDOW); (r, s, shift) track the tap
template <int KS, int MS>
__device__ __forceinline__ void issue_stage(bool leader, int nt, bool first_stage, uint32_t a_lo0, uint32_t b_lo0,
                                            uint32_t b_tap16, uint32_t row16, uint32_t jstride, uint32_t d0, uint32_t dstep,
                                            uint32_t desc_hi, uint32_t idesc, int kw, int row_w, int& s, int& shift_rows) {
  uint32_t b_lo = b_lo0;
  for (int tt = 0; tt < nt; ++tt) {
    const uint32_t a_lo = a_lo0 + (uint32_t)shift_rows * row16;
    issue_tap<KS, MS>(leader, a_lo, b_lo, jstride, d0, dstep, desc_hi, idesc, (first_stage && tt == 0) ? 0u : 1u);
    b_lo += b_tap16;
    ++shift_rows;
    if (++s == kw) { s = 0; shift_rows += row_w - kw; }
  }
}

template <bool WINDOW, bool FP16>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the swizzled operand tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int row_bytes = p.row_bytes;                                  // 128 or 64
  const int a_tile_bytes = kBlockM * row_bytes;                       // IM2COL stage A part
  const int b_tap_bytes = p.tile_n * row_bytes;
  const int b_stage_bytes = WINDOW ? p.b_group * b_tap_bytes : a_tile_bytes + b_tap_bytes;
  // IM2COL: `stages` x [A | B]          WINDOW: a_slots x A patch, then `stages` x (b_group B tiles)
  uint8_t* a_base = smem;
  uint8_t* b_base = WINDOW ? smem + (size_t)p.a_slots * p.a_slot_bytes : smem + a_tile_bytes;
  uint8_t* bar_base = WINDOW ? b_base + (size_t)p.stages * b_stage_bytes : smem + (size_t)p.stages * b_stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);     // [stages]  (B, or A+B for IM2COL)
  uint64_t* empty_bar = full_bar + kMaxStages;                    // [stages]
  uint64_t* a_full = full_bar + 2 * kMaxStages;                   // [a_slots] (WINDOW)
  uint64_t* a_empty = a_full + 4;
  uint64_t* tmem_full = a_full + 8;                               // [2]
  uint64_t* tmem_empty = a_full + 10;                             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 12);
  float* s_scale = reinterpret_cast<float*>(bar_base + kBarBytes);
  float* s_shift = s_scale + p.cout_pad;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_shift + p.cout_pad);      // 8 x 4 KB, 128-byte aligned
  for (int i = threadIdx.x; i < p.cout_pad; i += kThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }

  // warp index through a shuffle: the compiler then KNOWS it is warp-uniform, keeps the role
  // branches uniform and the producer / MMA loops on the uniform datapath (no R2UR per operand)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
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
  const int row_elems = row_bytes >> 1;
  const bool skip_loads = (p.debug_flags & 1) != 0;
  const bool skip_mma = (p.debug_flags & 2) != 0;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loops, one elected lane issues) =====================
    int stage = 0, aslot = 0;
    uint32_t phase = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
      const int n0 = n_tile * p.tile_n;
      if (WINDOW) {
        const int i0 = m_tile * tile_rows;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          ptx::mbar_wait(a_empty + aslot, aphase ^ 1);
          if (ptx::elect_one()) {
            uint8_t* dst = a_base + (size_t)aslot * p.a_slot_bytes;
            if (skip_loads) {
              ptx::mbar_arrive(a_full + aslot);
            } else {
              ptx::mbar_arrive_expect_tx(a_full + aslot, (uint32_t)(p.n_boxes * p.box_rows * row_bytes));
              for (int b = 0; b < p.n_boxes; ++b)
                ptx::tma_load_2d(dst + (size_t)b * p.box_rows * row_bytes, &tmap_a, a_full + aslot, cb * row_elems,
                                 i0 + b * p.box_rows);
            }
          }
          __syncwarp();
          if (++aslot == p.a_slots) { aslot = 0; aphase ^= 1; }
          for (int t0 = 0; t0 < taps; t0 += p.b_group) {
            const int nt = min(p.b_group, taps - t0);
            ptx::mbar_wait(empty_bar + stage, phase ^ 1);
            if (ptx::elect_one()) {
              if (skip_loads) {
                ptx::mbar_arrive(full_bar + stage);
              } else {
                uint8_t* dst = b_base + (size_t)stage * b_stage_bytes;
                ptx::mbar_arrive_expect_tx(full_bar + stage, (uint32_t)(nt * b_tap_bytes));
                for (int tt = 0; tt < nt; ++tt)
                  ptx::tma_load_2d(dst + (size_t)tt * b_tap_bytes, &tmap_b, full_bar + stage,
                                   ((t0 + tt) * p.cblocks + cb) * row_elems, n0);
              }
            }
            __syncwarp();
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
        int kcol = 0;
        for (int r = 0; r < p.kh; ++r) {
          for (int s = 0; s < p.kw; ++s) {
            for (int cb = 0; cb < p.cblocks; ++cb, kcol += row_elems) {
              ptx::mbar_wait(empty_bar + stage, phase ^ 1);
              if (ptx::elect_one()) {
                uint8_t* a_dst = smem + (size_t)stage * b_stage_bytes;
                if (skip_loads) {
                  ptx::mbar_arrive(full_bar + stage);
                } else {
                  ptx::mbar_arrive_expect_tx(full_bar + stage, (uint32_t)b_stage_bytes);
                  ptx::tma_load_im2col_4d(a_dst, &tmap_a, full_bar + stage, cb * row_elems, w0, h0, img,
                                          (uint16_t)s, (uint16_t)r);
                  ptx::tma_load_2d(a_dst + a_tile_bytes, &tmap_b, full_bar + stage, kcol, n0);
                }
              }
              __syncwarp();
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loops, one elected lane issues) =====================
    const uint32_t idesc = ptx::umma_idesc_f16(kBlockM, p.tile_n, FP16 ? 1 : 0);
    const uint32_t desc_hi = ptx::umma_desc_hi(row_bytes);
    const int full_ksteps = row_bytes >> 5;
    const bool leader = ptx::elect_one();      // deterministic: the same lane every time
    const uint32_t row16 = (uint32_t)row_bytes >> 4;                   // one operand row, in descriptor units
    const uint32_t b_tap16 = (uint32_t)b_tap_bytes >> 4;
    const uint32_t jstride = (uint32_t)(kBlockM * row_bytes) >> 4;     // next 128-row accumulator
    bool ready = false;
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
          const uint32_t a_addr = ptx::smem_u32(a_base + (size_t)aslot * p.a_slot_bytes);
          const int ksteps = (cb == p.cblocks - 1) ? p.last_ksteps : full_ksteps;
          int s = 0, shift_rows = p.win_shift0;
          const uint32_t a_lo0 = ptx::umma_desc_lo(a_addr);
          for (int t0 = 0; t0 < taps; t0 += p.b_group) {
            const int nt = min(p.b_group, taps - t0);
            ptx::mbar_wait(full_bar + stage, phase);
            ptx::tc_fence_after();
            const uint32_t b_lo0 = ptx::umma_desc_lo(ptx::smem_u32(b_base + (size_t)stage * b_stage_bytes));
            const bool first_stage = (cb == 0 && t0 == 0);
            if (!skip_mma) {
#define IFCB_ISSUE(KS, MS)                                                                                        \
  issue_stage<KS, MS>(leader, nt, first_stage, a_lo0, b_lo0, b_tap16, row16, jstride, d_tmem, (uint32_t)p.tile_n, \
                      desc_hi, idesc, p.kw, p.row_w, s, shift_rows)
              switch (ksteps * 8 + p.m_sub) {
                case 1 * 8 + 1: IFCB_ISSUE(1, 1); break;
                case 2 * 8 + 1: IFCB_ISSUE(2, 1); break;
                case 3 * 8 + 1: IFCB_ISSUE(3, 1); break;
                case 4 * 8 + 1: IFCB_ISSUE(4, 1); break;
                case 1 * 8 + 2: IFCB_ISSUE(1, 2); break;
                case 2 * 8 + 2: IFCB_ISSUE(2, 2); break;
                case 3 * 8 + 2: IFCB_ISSUE(3, 2); break;
                case 4 * 8 + 2: IFCB_ISSUE(4, 2); break;
                case 1 * 8 + 4: IFCB_ISSUE(1, 4); break;
                case 2 * 8 + 4: IFCB_ISSUE(2, 4); break;
                case 3 * 8 + 4: IFCB_ISSUE(3, 4); break;
                default: IFCB_ISSUE(4, 4); break;
              }
#undef IFCB_ISSUE
            }
            if (leader) ptx::umma_commit(empty_bar + stage);
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          if (leader) ptx::umma_commit(a_empty + aslot);
          __syncwarp();
          if (++aslot == p.a_slots) { aslot = 0; aphase ^= 1; }
        }
      } else {
        const int kblocks = taps * p.cblocks;
        int cb = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          // `ready` = a non-blocking peek taken one k-block earlier: its latency hides behind
          // the previous block's MMA issue
          if (!ready) ptx::mbar_wait(full_bar + stage, phase);
          ptx::tc_fence_after();
          const uint32_t a_lo = ptx::umma_desc_lo(ptx::smem_u32(smem + (size_t)stage * b_stage_bytes));
          const uint32_t b_lo = a_lo + (uint32_t)(a_tile_bytes >> 4);
          const int ksteps = (cb == p.cblocks - 1) ? p.last_ksteps : full_ksteps;
          const int cur = stage;
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          ready = ptx::mbar_test_wait(full_bar + stage, phase) != 0;
          const uint32_t first = kb > 0 ? 1u : 0u;
          if (!skip_mma) {
            switch (ksteps) {
              case 1: issue_tap<1, 1>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
              case 2: issue_tap<2, 1>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
              case 3: issue_tap<3, 1>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
              default: issue_tap<4, 1>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
            }
          }
          if (leader) ptx::umma_commit(empty_bar + cur);     // frees the smem slot when these MMAs retire
          __syncwarp();
          if (++cb == p.cblocks) cb = 0;
        }
      }
      if (leader) ptx::umma_commit(tmem_full + acc);           // accumulators ready for the epilogue
      __syncwarp();
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
      int gcount = 0;
      for (int j = 0; j < p.m_sub && !(p.debug_flags & 4); ++j) {
        const int row0 = m_tile * tile_rows + j * kBlockM + quad * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kAccBufCols + j * p.tile_n);
        epilogue_128rows<FP16>(p, taddr, row0, n_tile * p.tile_n, lane, half, gcount, stage_addr, s_scale, s_shift);
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

// Shared-memory plan of a layer (kp.row_bytes, tile_n, cout_pad, kh, kw set).  Returns false if nothing fits.
bool conv_plan_smem(ConvKernelParams& kp, bool window, int halo_rows) {
  const int row_bytes = kp.row_bytes;
  const int b_tap = kp.tile_n * row_bytes;
  const int a_tile = kBlockM * row_bytes;
  const int fixed = epilogue_smem(kp.cout_pad) + 1024;
  const int taps = kp.kh * kp.kw;
  if (!window) {
    kp.m_sub = 1;
    kp.a_slots = 0;
    kp.a_slot_bytes = 0;
    kp.box_rows = kp.n_boxes = 0;
    kp.b_group = 1;
    int s = (kSmemBudget - fixed) / (a_tile + b_tap);
    if (s > 10) s = 10;
    kp.stages = s;
    return s >= 2;
  }
  // taps per B pipeline stage: about 24 KB per stage, balanced over the groups
  int g = 24576 / b_tap;
  if (g < 1) g = 1;
  if (g > taps) g = taps;
  if (kp.b_group_cap > 0 && g > kp.b_group_cap) g = kp.b_group_cap;
  const int n_groups = (taps + g - 1) / g;
  g = (taps + n_groups - 1) / n_groups;
  const int b_stage = g * b_tap;
  // choose m (accumulators per tile) as large as TMEM double buffering and smem allow
  for (int m = 4; m >= 1; m >>= 1) {
    if (m > kp.m_sub_cap) continue;
    if (m * kp.tile_n > kAccBufCols) continue;
    const int rows = kBlockM * m + halo_rows;
    const int n_boxes = (rows + 255) / 256;
    int box_rows = ((rows + n_boxes - 1) / n_boxes + 7) & ~7;
    const int slot = (n_boxes * box_rows * row_bytes + 1023) & ~1023;
    for (int slots = kp.a_slots_pref; slots >= 1; --slots) {
      const int left = kSmemBudget - fixed - slots * slot;
      int s = left / b_stage;
      if (s > 8) s = 8;
      const int need = (g >= 2 || slots >= 2) ? 2 : 3;
      if (s >= need) {
        kp.m_sub = m;
        kp.a_slots = slots;
        kp.a_slot_bytes = slot;
        kp.box_rows = box_rows;
        kp.n_boxes = n_boxes;
        kp.stages = s;
        kp.b_group = g;
        return true;
      }
      if (m > 1) break;      // prefer a smaller m with two slots over a single-slot large m
    }
  }
  return false;
}

int conv_smem_bytes(const ConvKernelParams& kp, bool window) {
  const int b_tap = kp.tile_n * kp.row_bytes;
  const int ops = window ? kp.a_slots * kp.a_slot_bytes + kp.stages * kp.b_group * b_tap
                         : kp.stages * (kBlockM * kp.row_bytes + b_tap);
  return ops + epilogue_smem(kp.cout_pad) + 1024;
}

namespace {
template <bool WINDOW, bool FP16>
int launch_variant(const ConvLayer& L, const ConvKernelParams& p, int grid, int smem, cudaStream_t stream) {
  static int attr_smem = 0;
  if (smem > attr_smem) {
    IFCB_CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<WINDOW, FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  conv_umma_kernel<WINDOW, FP16><<<grid, kThreads, smem, stream>>>(L.tmap_a, L.tmap_b, p);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
}  // namespace

int launch_conv(const ConvLayer& L, int batch, cudaStream_t stream) {
  ConvKernelParams p = L.kp;
  p.rows = (long long)batch * p.rows_per_img;
  const int tile_rows = kBlockM * p.m_sub;
  const long long m_tiles = (p.rows + tile_rows - 1) / tile_rows;
  const long long total = m_tiles * p.n_tiles;
  if (total == 0) return 0;
  if (p.rows + tile_rows >= (1ll << 31)) {
    set_error("conv: %lld GEMM rows exceed the 32-bit row index", p.rows);
    return -1;
  }
  const int grid = (int)(total < sm_count() ? total : sm_count());
  const int smem = conv_smem_bytes(p, L.window);
  if (L.window) return p.fp16 ? launch_variant<true, true>(L, p, grid, smem, stream) : launch_variant<true, false>(L, p, grid, smem, stream);
  return p.fp16 ? launch_variant<false, true>(L, p, grid, smem, stream) : launch_variant<false, false>(L, p, grid, smem, stream);
}

// ---------------------------------------------------------------------------------
// Debug probe (tests only): issue ONE im2col TMA load with a layer's tensor map
// and copy the raw shared-memory tile (still swizzled) to global, so
// the TMA im2col semantics can be checked independently of the MMA path.
// ---------------------------------------------------------------------------------
namespace {
__global__ void im2col_probe_kernel(const __grid_constant__ CUtensorMap tmap_a, int c, int w, int h, int n,
                                    int off_w, int off_h, int tile_bytes, uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + tile_bytes);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar, tile_bytes);
    ptx::tma_load_im2col_4d(smem, &tmap_a, bar, c, w, h, n, (uint16_t)off_w, (uint16_t)off_h);
  }
  ptx::mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < tile_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(smem)[i];
}
}  // namespace

int launch_im2col_probe(const ConvLayer& L, int c, int w, int h, int n, int off_w, int off_h, void* d_out,
                        cudaStream_t stream) {
  const int tile_bytes = kBlockM * L.kp.row_bytes;
  const int smem = tile_bytes + 1024 + 64;
  im2col_probe_kernel<<<1, 128, smem, stream>>>(L.tmap_a, c, w, h, n, off_w, off_h, tile_bytes,
                                                reinterpret_cast<uint8_t*>(d_out));
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace ifcb
