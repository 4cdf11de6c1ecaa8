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
//    epilogue drops.  Layers with >= 64 output channels run this scheme on CTA PAIRS
//    (conv_umma_pair_kernel<WINDOW = true>): the shared-memory port bounds these layers, and a
//    pair halves the weight bytes every SM writes and keeps.
//
// Operand rows are one swizzle span wide: 128 bytes (64 channels, SWIZZLE_128B) or, for
// Cin <= 32, 64 bytes (32 channels, SWIZZLE_64B) -- halves the shared-memory footprint of
// the first layers.
//
// Persistent, warp-specialised CTA (1 per SM, 448 threads):
//   warp 0      TMA producer
//   warp 1      TMEM allocator + tcgen05.mma issuer
//   warps 2-13  epilogue: tcgen05.ld -> scale/shift (+residual) -> ReLU -> 16-bit ->
//               swizzled smem tile -> coalesced 16-byte stores (three warps per TMEM lane
//               quadrant, taking (32-column slice, accumulator) units round-robin)
// TRAIN (STATS kernels): the epilogue also accumulates, per output channel, the sum and the sum of
// squares of the 16-bit values it stores -- BatchNorm's batch statistics without a pass over z.
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
#ifndef IFCB_EPI_WARPS
#define IFCB_EPI_WARPS 12
#endif
constexpr int kEpiWarps = IFCB_EPI_WARPS;           // 3-4 per TMEM lane quadrant: the epilogue is latency-bound, TLP hides it
constexpr int kEpiPerQuad = kEpiWarps / 4;
constexpr int kThreads = 64 + 32 * kEpiWarps;       // 448
constexpr int kStageTile = 2048;                    // per-warp staging tile: 32 rows x 64 B
constexpr int kTmemCols = 512;
constexpr int kAccBufCols = 256;
constexpr int kBarBytes = 512;
constexpr int kMaxStages = 12;

// barriers + scale/shift (2 x cout_pad floats) + one 2 KB staging tile per epilogue warp
// (+ per-channel sum / sum-of-squares floats when the epilogue gathers BatchNorm statistics)
__host__ __device__ constexpr int epilogue_smem(int cout_pad, bool stats = false) {
  return kBarBytes + 8 * cout_pad + kEpiWarps * kStageTile + (stats ? 8 * cout_pad : 0);
}

__device__ __forceinline__ void decode_tile(const ConvKernelParams& p, int tile, int m_tiles, int& m_tile, int& n_tile) {
  if (p.n_major) {
    n_tile = tile / m_tiles;
    m_tile = tile - n_tile * m_tiles;
  } else {
    m_tile = tile / p.n_tiles;
    n_tile = tile - m_tile * p.n_tiles;
  }
}

struct RowMap {           // where a GEMM row lands
  int n, p, q;
  bool valid;
};

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

// y = a * b + c on two packed fp32 lanes (sm_100 FFMA2): halves the epilogue's FMA count
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n.reg .b64 a, b, c, d;\nmov.b64 a, {%2, %3};\nmov.b64 b, {%4, %5};\nmov.b64 c, {%6, %7};\n"
      "fma.rn.f32x2 d, a, b, c;\nmov.b64 {%0, %1}, d;\n}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}

constexpr uint32_t kNoRow = 0xFFFFFFFFu;

// coalesced write-out of one staged unit: 32 rows x PPR 16-byte pieces (row pitch 64 B, chunk
// XOR-swizzled by (row >> 1) & 3), fully unrolled.  `off` = this lane's destination row as an
// ELEMENT offset (row * ld), kNoRow = dropped row.
// STATS: this lane's PPR rows x 8 channels (the 16-bit values that go to memory) are also added into
// acc[0..8) (sums) / acc[8..16) (sums of squares); dropped rows are skipped.
template <int PPR, bool FP16, bool STATS>
__device__ __forceinline__ void write_out(const uint8_t* stage, int lane, uint32_t off, __nv_bfloat16* gout, float (&acc)[16]) {
  uint4 val[PPR];
  uint32_t ro[PPR];
#pragma unroll
  for (int it = 0; it < PPR; ++it) {
    const int idx = lane + 32 * it;
    const int r = PPR == 4 ? idx >> 2 : idx >> 1;
    const int pc = idx - r * PPR;
    ro[it] = __shfl_sync(0xffffffffu, off, r);
    val[it] = *reinterpret_cast<const uint4*>(stage + r * 64 + ((pc ^ ((r >> 1) & 3)) << 4));
  }
  if (STATS) {
#pragma unroll
    for (int it = 0; it < PPR; ++it) {
      if (ro[it] != kNoRow) {
        const uint32_t w[4] = {val[it].x, val[it].y, val[it].z, val[it].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_act2(w[e], FP16 ? 1 : 0);
          acc[2 * e] += f.x;
          acc[2 * e + 1] += f.y;
          acc[8 + 2 * e] = fmaf(f.x, f.x, acc[8 + 2 * e]);
          acc[8 + 2 * e + 1] = fmaf(f.y, f.y, acc[8 + 2 * e + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int it = 0; it < PPR; ++it) {
    const int idx = lane + 32 * it;
    const int r = PPR == 4 ? idx >> 2 : idx >> 1;
    const int pc = idx - r * PPR;
    if (ro[it] != kNoRow) *reinterpret_cast<uint4*>(gout + ro[it] + pc * 8) = val[it];
  }
}

// ---------------------------------------------------------------------------------------
// Epilogue of one tile (m_sub accumulators of 128 rows x tile_n columns), executed by the four
// warps of one TMEM lane quadrant.  Work unit = (32-column slice, accumulator); the four warps
// take units round-robin.  Row mapping (two magic divisions) once per accumulator and tile;
// destination rows travel as 32-bit element offsets so that the write-out loop is
// shuffle + ld.shared + one wide multiply-add + store.
// ---------------------------------------------------------------------------------------
template <bool FP16, bool STATS>
__device__ __forceinline__ void epilogue_tile(const ConvKernelParams& p, uint32_t tmem_acc, int tile_row0, int n_lo, int lane,
                                              int sub, uint8_t* stage, const float* s_scale, const float* s_shift,
                                              float (&sacc)[3][16]) {
  const int n_hi = n_lo + p.tile_n;
  const int ms_shift = p.m_sub == 1 ? 0 : p.m_sub == 2 ? 1 : 2;
  uint4* srow = reinterpret_cast<uint4*>(stage + lane * 64);
  const int sw = (lane >> 1) & 3;
  int next = sub;                                    // tile-wide index of this warp's next unit
  int ubase = 0;                                     // tile-wide index of the segment's first unit
  int slot = 0;                                      // STATS: how many units this warp has done in this tile (<= 3: one segment)
  for (int si = 0; si < p.n_seg; ++si) {
    const int g_lo = max(n_lo, p.seg_begin[si]), g_hi = min(n_hi, p.seg_end[si]);
    if (g_lo >= g_hi) continue;
    const int nunits = ((g_hi - g_lo + 31) >> 5) << ms_shift;
    if (next >= ubase + nunits) { ubase += nunits; continue; }     // none of this segment's units is ours
    const int relu = p.seg_relu[si];
    const uint32_t ld = (uint32_t)p.seg_ld[si];
    const int ph = p.seg_pad_h[si], pw = p.seg_pad_w[si];
    const int Hd = p.P + 2 * ph, Wd = p.Q + 2 * pw;
    __nv_bfloat16* seg_out = p.seg_out[si] + (g_lo - p.seg_begin[si]);
    for (; next < ubase + nunits; next += kEpiPerQuad) {
      const int u = next - ubase;
      const int gi = u >> ms_shift;
      const int j = u - (gi << ms_shift);
      const int g0 = g_lo + gi * 32;
      const bool wide = (g_hi - g0) >= 32;                 // 32 columns, or a 16-column tail
      // where does this lane's row of accumulator j land?  element offsets, kNoRow = dropped
      uint32_t off_j = kNoRow, roff_j = kNoRow;
      {
        const int i = tile_row0 + j * kBlockM + lane;
        if (p.identity_rows) {
          if (i < (int)p.rows) {
            off_j = (uint32_t)i * ld;
            roff_j = (uint32_t)i * (uint32_t)p.res_ld;
          }
        } else {
          const RowMap rm = map_row(p, i);
          if (rm.valid) {
            off_j = (uint32_t)((rm.n * Hd + rm.p + ph) * Wd + rm.q + pw) * ld;
            roff_j = (uint32_t)((rm.n * (p.P + 2 * p.res_pad_h) + rm.p + p.res_pad_h) * (p.Q + 2 * p.res_pad_w) + rm.q + p.res_pad_w) * (uint32_t)p.res_ld;
          }
        }
      }
      const uint32_t taddr = tmem_acc + (uint32_t)(j * p.tile_n + g0 - n_lo);
      uint32_t v[2][16];
      ptx::tmem_ld_32x32b_x16(taddr, v[0]);
      if (wide) ptx::tmem_ld_32x32b_x16(taddr + 16u, v[1]);
      ptx::tmem_ld_wait();
      uint4 o[4];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        if (ch == 0 || wide) {
          const int n = g0 + ch * 16;
          float y[16];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + n + 4 * q4);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + n + 4 * q4);
            ffma2(y[4 * q4 + 0], y[4 * q4 + 1], __uint_as_float(v[ch][4 * q4 + 0]), __uint_as_float(v[ch][4 * q4 + 1]),
                  sc.x, sc.y, sh.x, sh.y);
            ffma2(y[4 * q4 + 2], y[4 * q4 + 3], __uint_as_float(v[ch][4 * q4 + 2]), __uint_as_float(v[ch][4 * q4 + 3]),
                  sc.z, sc.w, sh.z, sh.w);
          }
          if (p.residual != nullptr && roff_j != kNoRow) {
            const uint4* rp4 = reinterpret_cast<const uint4*>(p.residual + roff_j + n);
            const uint4 r0 = __ldg(rp4), r1 = __ldg(rp4 + 1);
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float2 f = unpack_act2(rr[e], FP16 ? 1 : 0);
              y[2 * e] += f.x;
              y[2 * e + 1] += f.y;
            }
          }
          if (relu) pack16<FP16, true>(y, o[2 * ch], o[2 * ch + 1]);
          else pack16<FP16, false>(y, o[2 * ch], o[2 * ch + 1]);
        }
      }
      // stage: row `lane` (64 B), 16-byte pieces XOR-swizzled by (row >> 1) & 3
      srow[0 ^ sw] = o[0];
      srow[1 ^ sw] = o[1];
      if (wide) {
        srow[2 ^ sw] = o[2];
        srow[3 ^ sw] = o[3];
      }
      __syncwarp();
      if (STATS) {
        // the unit -> slot assignment is the same in every tile of a column tile, so the per-lane partial sums stay in
        // registers across tiles (stats_flush reduces them over the lanes once per column tile)
#define IFCB_WRITE_OUT(SLOT)                                                                   \
  do {                                                                                         \
    if (wide) write_out<4, FP16, true>(stage, lane, off_j, seg_out + gi * 32, sacc[SLOT]);     \
    else write_out<2, FP16, true>(stage, lane, off_j, seg_out + gi * 32, sacc[SLOT]);          \
  } while (0)
        if (slot == 0) IFCB_WRITE_OUT(0);
        else if (slot == 1) IFCB_WRITE_OUT(1);
        else IFCB_WRITE_OUT(2);
#undef IFCB_WRITE_OUT
        ++slot;
      } else {
        if (wide) write_out<4, FP16, false>(stage, lane, off_j, seg_out + gi * 32, sacc[0]);
        else write_out<2, FP16, false>(stage, lane, off_j, seg_out + gi * 32, sacc[0]);
      }
      __syncwarp();
    }
    ubase += nunits;
  }
}

// Reduces a warp's per-lane statistics partials over the lanes and adds them into the CTA's shared per-channel sums; clears
// the partials.  Mirrors epilogue_tile's unit walk for ONE segment: unit u = sub + 3 * slot covers columns g0 .. g0 + 32 (or a
// 16-column tail); a lane's partials belong to the 8 channels of its 16-byte piece pc = lane & 3 (wide) / lane & 1 (tail).
__device__ __forceinline__ void stats_flush(const ConvKernelParams& p, float (&sacc)[3][16], int n_lo, int lane, int sub,
                                            float* s_sum, float* s_sq) {
  const int n_hi = n_lo + p.tile_n;
  const int ms_shift = p.m_sub == 1 ? 0 : p.m_sub == 2 ? 1 : 2;
  const int g_lo = max(n_lo, p.seg_begin[0]), g_hi = min(n_hi, p.seg_end[0]);
  const int nunits = g_lo < g_hi ? (((g_hi - g_lo + 31) >> 5) << ms_shift) : 0;
#pragma unroll
  for (int slot = 0; slot < 3; ++slot) {
    const int u = sub + kEpiPerQuad * slot;
    if (u < nunits) {
      const int g0 = g_lo + (u >> ms_shift) * 32;
      const bool wide = (g_hi - g0) >= 32;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float v = sacc[slot][j];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (!wide) v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (lane < (wide ? 4 : 2)) atomicAdd((j < 8 ? s_sum : s_sq) + g0 + lane * 8 + (j & 7), v);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) sacc[slot][j] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------
// MMA issue, specialised on the K steps per operand row (KS) and the accumulators per tile
// (MS) so that every tcgen05.mma of a filter tap is straight-line code on the uniform
// datapath: one tap = MS x KS MMAs, operands a_lo + j*jstride + 2k / b_lo + 2k.
// ---------------------------------------------------------------------------------------
template <int KS, int MS, bool PAIR = false>
__device__ __forceinline__ void issue_tap(bool leader, uint32_t a_lo, uint32_t b_lo, uint32_t jstride, uint32_t d0,
                                          uint32_t dstep, uint32_t desc_hi, uint32_t idesc, uint32_t first) {
#pragma unroll
  for (int j = 0; j < MS; ++j) {
#pragma unroll
    for (int k = 0; k < KS; ++k)
      if (leader) {
        if (PAIR)
          ptx::umma_f16_lohi_pair(d0 + (uint32_t)j * dstep, a_lo + (uint32_t)j * jstride + (uint32_t)(2 * k),
                                  b_lo + (uint32_t)(2 * k), desc_hi, idesc, k > 0 ? 1u : first);
        else
          ptx::umma_f16_lohi(d0 + (uint32_t)j * dstep, a_lo + (uint32_t)j * jstride + (uint32_t)(2 * k),
                             b_lo + (uint32_t)(2 * k), desc_hi, idesc, k > 0 ? 1u : first);
      }
  }
}

// all taps [t0, t0 + nt) of one weight pipeline stage (WINDOW); (r, s, shift) track the tap
template <int KS, int MS, bool PAIR = false>
__device__ __forceinline__ void issue_stage(bool leader, int nt, bool first_stage, uint32_t a_lo0, uint32_t b_lo0,
                                            uint32_t b_tap16, uint32_t row16, uint32_t jstride, uint32_t d0, uint32_t dstep,
                                            uint32_t desc_hi, uint32_t idesc, int kw, int row_w, int& s, int& shift_rows) {
  uint32_t b_lo = b_lo0;
  for (int tt = 0; tt < nt; ++tt) {
    const uint32_t a_lo = a_lo0 + (uint32_t)shift_rows * row16;
    issue_tap<KS, MS, PAIR>(leader, a_lo, b_lo, jstride, d0, dstep, desc_hi, idesc, (first_stage && tt == 0) ? 0u : 1u);
    b_lo += b_tap16;
    ++shift_rows;
    if (++s == kw) { s = 0; shift_rows += row_w - kw; }
  }
}

template <bool WINDOW, bool FP16, bool STATS>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the swizzled operand tiles
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (keeps the address space known: LDS/STS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int row_bytes = p.row_bytes;                                  // 128 or 64
  const int a_tile_bytes = kBlockM * row_bytes;                       // IM2COL stage A part
  const int b_tap_bytes = p.tile_n * row_bytes;
  const int kb_bytes = a_tile_bytes + b_tap_bytes;                    // IM2COL: one k-block (A tile + B tile)
  const int b_stage_bytes = WINDOW ? p.b_group * b_tap_bytes : p.b_group * kb_bytes;
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
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_shift + p.cout_pad);      // 12 x 2 KB, 128-byte aligned
  float* s_sum = reinterpret_cast<float*>(s_stage + kEpiWarps * kStageTile);  // STATS: [cout_pad] sums, [cout_pad] sums of squares
  float* s_sq = s_sum + p.cout_pad;
  for (int i = threadIdx.x; i < p.cout_pad; i += kThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
    if (STATS) { s_sum[i] = 0.f; s_sq[i] = 0.f; }
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
      int m_tile, n_tile;
      decode_tile(p, tile, m_tiles, m_tile, n_tile);
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
        // k-blocks (filter tap x channel block) travel in groups of p.b_group per pipeline stage:
        // one barrier round trip per group
        const int kblocks = taps * p.cblocks;
        int kcol = 0, r = 0, s = 0, cb = 0;
        for (int kb = 0; kb < kblocks; kb += p.b_group) {
          const int nk = min(p.b_group, kblocks - kb);
          ptx::mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* dst = smem + (size_t)stage * b_stage_bytes;
          const bool el = ptx::elect_one();
          if (el) {
            if (skip_loads) ptx::mbar_arrive(full_bar + stage);
            else ptx::mbar_arrive_expect_tx(full_bar + stage, (uint32_t)(nk * kb_bytes));
          }
          for (int i = 0; i < nk; ++i) {
            if (el && !skip_loads) {
              ptx::tma_load_im2col_4d(dst + (size_t)i * kb_bytes, &tmap_a, full_bar + stage, cb * row_elems, w0, h0, img,
                                      (uint16_t)s, (uint16_t)r);
              ptx::tma_load_2d(dst + (size_t)i * kb_bytes + a_tile_bytes, &tmap_b, full_bar + stage, kcol, n0);
            }
            kcol += row_elems;
            if (++cb == p.cblocks) { cb = 0; if (++s == p.kw) { s = 0; ++r; } }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
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
            if (!ready) ptx::mbar_wait(full_bar + stage, phase);
            ptx::tc_fence_after();
            const uint32_t b_lo0 = ptx::umma_desc_lo(ptx::smem_u32(b_base + (size_t)stage * b_stage_bytes));
            const bool first_stage = (cb == 0 && t0 == 0);
            const int cur = stage;
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            ready = ptx::mbar_test_wait(full_bar + stage, phase) != 0;     // early peek at the next stage
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
            if (leader) ptx::umma_commit(empty_bar + cur);
            __syncwarp();
          }
          if (leader) ptx::umma_commit(a_empty + aslot);
          __syncwarp();
          if (++aslot == p.a_slots) { aslot = 0; aphase ^= 1; }
        }
      } else {
        const int kblocks = taps * p.cblocks;
        const uint32_t kb16 = (uint32_t)kb_bytes >> 4;
        int cb = 0;
        for (int kb = 0; kb < kblocks; kb += p.b_group) {
          const int nk = min(p.b_group, kblocks - kb);
          // `ready` = a non-blocking peek taken one stage earlier: its latency hides behind
          // the previous stage's MMA issue
          if (!ready) ptx::mbar_wait(full_bar + stage, phase);
          ptx::tc_fence_after();
          uint32_t a_lo = ptx::umma_desc_lo(ptx::smem_u32(smem + (size_t)stage * b_stage_bytes));
          const int cur = stage;
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          ready = ptx::mbar_test_wait(full_bar + stage, phase) != 0;
          for (int i = 0; i < nk; ++i, a_lo += kb16) {
            const uint32_t b_lo = a_lo + (uint32_t)(a_tile_bytes >> 4);
            const int ksteps = (cb == p.cblocks - 1) ? p.last_ksteps : full_ksteps;
            const uint32_t first = (kb + i) > 0 ? 1u : 0u;
            if (!skip_mma) {
              switch (ksteps) {
                case 1: issue_tap<1, 1>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
                case 2: issue_tap<2, 1>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
                case 3: issue_tap<3, 1>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
                default: issue_tap<4, 1>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
              }
            }
            if (++cb == p.cblocks) cb = 0;
          }
          if (leader) ptx::umma_commit(empty_bar + cur);     // frees the smem slot when these MMAs retire
          __syncwarp();
        }
      }
      if (leader) ptx::umma_commit(tmem_full + acc);           // accumulators ready for the epilogue
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int sub = (warp - 2) >> 2;                 // which of the quadrant's four warps
    uint8_t* stage = s_stage + (warp - 2) * kStageTile;
    float sacc[STATS ? 3 : 1][16];
    if (STATS) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < 16; ++j) sacc[STATS ? a : 0][j] = 0.f;
    }
    int cur_n = -1;
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      int m_tile, n_tile;
      decode_tile(p, tile, m_tiles, m_tile, n_tile);
      if (STATS && n_tile != cur_n) {                 // the column tile changes: hand the finished one's partial sums over
        if (cur_n >= 0) stats_flush(p, reinterpret_cast<float(&)[3][16]>(sacc), cur_n * p.tile_n, lane, sub, s_sum, s_sq);
        cur_n = n_tile;
      }
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      ptx::mbar_wait(tmem_full + acc, acc_phase);
      ptx::tc_fence_after();
      if (!(p.debug_flags & 4)) {
        const uint32_t tmem_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kAccBufCols);
        epilogue_tile<FP16, STATS>(p, tmem_acc, m_tile * tile_rows + quad * 32, n_tile * p.tile_n, lane, sub, stage, s_scale, s_shift,
                                   reinterpret_cast<float(&)[3][16]>(sacc));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty + acc);
    }
    if (STATS && cur_n >= 0) stats_flush(p, reinterpret_cast<float(&)[3][16]>(sacc), cur_n * p.tile_n, lane, sub, s_sum, s_sq);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
  if (STATS) {            // one float64 atomic per channel, quantity and CTA
    for (int i = threadIdx.x; i < p.cout; i += kThreads) {
      atomicAdd(p.stats + i, (double)s_sum[i]);
      atomicAdd(p.stats + p.cout + i, (double)s_sq[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// IM2COL on CTA PAIRS (cluster of 2, tcgen05 cta_group::2): one tile = 256 output pixels x
// tile_n.  Each CTA loads its own 128 pixel rows of A and HALF of the weight tile; the leader
// CTA issues M = 256 MMAs that drive both SMs' tensor cores.  Per SM and k-block this halves
// the weight bytes both fetched from L2 and read from shared memory -- a 128-row tile streams
// weights at 64 B per MMA clock, which is the L2 -> SM limit (profiles/r01_layer_matrix_*.txt).
//   full[s]   leader only: expect_tx covers both CTAs' loads (peer TMA signals it remotely)
//   empty[s], tmem_full[a]   in each CTA, armed by multicast tcgen05.commit
//   tmem_empty[a]   leader only: every epilogue warp of both CTAs arrives on it
// ---------------------------------------------------------------------------------------
// WINDOW on CTA pairs (same kernel, WINDOW = true): each CTA loads the input patch of ITS m accumulators' rows and half of every
// weight tile (per SM half the weight bytes written to and kept in shared memory -- the shared-memory port is what bounds these
// layers); the leader's M = 256 MMAs take rows [0, 128) of accumulator j from the leader's patch and rows [128, 256) from the
// peer's patch at the same offset, so every tap is the same shifted descriptor as in the single-CTA kernel.
//   a_full[s]   leader only, like full[s]: both CTAs' patch loads signal it;   a_empty[s]   in each CTA (multicast commit)
template <bool WINDOW, bool FP16, bool STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_umma_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (keeps the address space known: LDS/STS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int row_bytes = p.row_bytes;
  const int a_tile_bytes = kBlockM * row_bytes;
  const int b_half_bytes = (p.tile_n >> 1) * row_bytes;
  const int kb_bytes = a_tile_bytes + b_half_bytes;                   // IM2COL: one k-block of this CTA
  // IM2COL: `stages` x b_group x [A | B half]        WINDOW: a_slots x A patch, then `stages` x (b_group B half tiles)
  const int stage_bytes = WINDOW ? p.b_group * b_half_bytes : p.b_group * kb_bytes;
  uint8_t* a_base = smem;
  uint8_t* b_base = WINDOW ? smem + (size_t)p.a_slots * p.a_slot_bytes : smem;
  uint8_t* bar_base = b_base + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* a_full = full_bar + 2 * kMaxStages;                       // [a_slots] (WINDOW)
  uint64_t* a_empty = a_full + 4;
  uint64_t* tmem_full = full_bar + 2 * kMaxStages + 8;
  uint64_t* tmem_empty = full_bar + 2 * kMaxStages + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 2 * kMaxStages + 12);
  float* s_scale = reinterpret_cast<float*>(bar_base + kBarBytes);
  float* s_shift = s_scale + p.cout_pad;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_shift + p.cout_pad);
  float* s_sum = reinterpret_cast<float*>(s_stage + kEpiWarps * kStageTile);
  float* s_sq = s_sum + p.cout_pad;
  for (int i = threadIdx.x; i < p.cout_pad; i += kThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
    if (STATS) { s_sum[i] = 0.f; s_sq[i] = 0.f; }
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar + s, 1);
      ptx::mbar_init(empty_bar + s, 1);
    }
    if (WINDOW)
      for (int s = 0; s < p.a_slots; ++s) {
        ptx::mbar_init(a_full + s, 1);
        ptx::mbar_init(a_empty + s, 1);
      }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tmem_full + s, 1);
      ptx::mbar_init(tmem_empty + s, 2 * kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, kTmemCols);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int tile_rows = kBlockM * p.m_sub;                            // rows of ONE CTA per tile (m_sub = 1 for IM2COL)
  const int m_tiles = (int)((p.rows + 2 * tile_rows - 1) / (2 * tile_rows));
  const int total_tiles = m_tiles * p.n_tiles;
  const int taps = p.kh * p.kw;
  const int row_elems = row_bytes >> 1;
  const bool skip_loads = (p.debug_flags & 1) != 0;
  const bool skip_mma = (p.debug_flags & 2) != 0;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    int aslot = 0;
    uint32_t aphase = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      int m_tile, n_tile;
      decode_tile(p, tile, m_tiles, m_tile, n_tile);
      const int n0 = n_tile * p.tile_n + (int)rank * (p.tile_n >> 1);
      if (WINDOW) {
        const int i0 = (m_tile * 2 + (int)rank) * tile_rows;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          ptx::mbar_wait(a_empty + aslot, aphase ^ 1);
          const uint32_t lead_afull = ptx::mapa(ptx::smem_u32(a_full + aslot), 0u);
          const bool el = ptx::elect_one();
          if (el) {
            uint8_t* dst = a_base + (size_t)aslot * p.a_slot_bytes;
            if (rank == 0) {
              if (skip_loads) ptx::mbar_arrive(a_full + aslot);
              else ptx::mbar_arrive_expect_tx(a_full + aslot, (uint32_t)(2 * p.n_boxes * p.box_rows * row_bytes));
            }
            if (!skip_loads)
              for (int b = 0; b < p.n_boxes; ++b)
                ptx::tma_load_2d_pair(dst + (size_t)b * p.box_rows * row_bytes, &tmap_a, lead_afull, cb * row_elems, i0 + b * p.box_rows);
          }
          __syncwarp();
          if (++aslot == p.a_slots) { aslot = 0; aphase ^= 1; }
          for (int t0 = 0; t0 < taps; t0 += p.b_group) {
            const int nt = min(p.b_group, taps - t0);
            ptx::mbar_wait(empty_bar + stage, phase ^ 1);
            const uint32_t lead_full = ptx::mapa(ptx::smem_u32(full_bar + stage), 0u);
            if (ptx::elect_one()) {
              uint8_t* dst = b_base + (size_t)stage * stage_bytes;
              if (rank == 0) {
                if (skip_loads) ptx::mbar_arrive(full_bar + stage);
                else ptx::mbar_arrive_expect_tx(full_bar + stage, (uint32_t)(2 * nt * b_half_bytes));
              }
              if (!skip_loads)
                for (int tt = 0; tt < nt; ++tt)
                  ptx::tma_load_2d_pair(dst + (size_t)tt * b_half_bytes, &tmap_b, lead_full, ((t0 + tt) * p.cblocks + cb) * row_elems, n0);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        continue;
      }
      const int m0 = m_tile * 2 * kBlockM + (int)rank * kBlockM;
      const int img = m0 / p.rows_per_img;
      const int rem = m0 - img * p.rows_per_img;
      const int op = rem / p.row_w, oq = rem - op * p.row_w;
      const int w0 = oq * p.stride_w - p.pad_w;
      const int h0 = op * p.stride_h - p.pad_h;
      const int kblocks = taps * p.cblocks;
      int kcol = 0, r = 0, s = 0, cb = 0;
      for (int kb = 0; kb < kblocks; kb += p.b_group) {
        const int nk = min(p.b_group, kblocks - kb);
        ptx::mbar_wait(empty_bar + stage, phase ^ 1);
        uint8_t* dst = smem + (size_t)stage * stage_bytes;
        const uint32_t lead_full = ptx::mapa(ptx::smem_u32(full_bar + stage), 0u);
        const bool el = ptx::elect_one();
        if (el && rank == 0) {
          if (skip_loads) ptx::mbar_arrive(full_bar + stage);
          else ptx::mbar_arrive_expect_tx(full_bar + stage, (uint32_t)(2 * nk * kb_bytes));
        }
        for (int i = 0; i < nk; ++i) {
          if (el && !skip_loads) {
            ptx::tma_load_im2col_4d_pair(dst + (size_t)i * kb_bytes, &tmap_a, lead_full, cb * row_elems, w0, h0, img, (uint16_t)s,
                                         (uint16_t)r);
            ptx::tma_load_2d_pair(dst + (size_t)i * kb_bytes + a_tile_bytes, &tmap_b, lead_full, kcol, n0);
          }
          kcol += row_elems;
          if (++cb == p.cblocks) { cb = 0; if (++s == p.kw) { s = 0; ++r; } }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      const uint32_t idesc = ptx::umma_idesc_f16(2 * kBlockM, p.tile_n, FP16 ? 1 : 0);
      const uint32_t desc_hi = ptx::umma_desc_hi(row_bytes);
      const int full_ksteps = row_bytes >> 5;
      const bool leader = ptx::elect_one();
      bool ready = false;
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      const int kblocks = taps * p.cblocks;
      const uint32_t kb16 = (uint32_t)kb_bytes >> 4;
      int aslot = 0;
      uint32_t aphase = 0;
      const uint32_t row16 = (uint32_t)row_bytes >> 4;
      const uint32_t b_tap16 = (uint32_t)b_half_bytes >> 4;
      const uint32_t jstride = (uint32_t)(kBlockM * row_bytes) >> 4;
      for (int tile = pair; tile < total_tiles; tile += npairs, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        ptx::mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccBufCols);
        if (WINDOW) {
          for (int cb = 0; cb < p.cblocks; ++cb) {
            ptx::mbar_wait(a_full + aslot, aphase);
            const uint32_t a_lo0 = ptx::umma_desc_lo(ptx::smem_u32(a_base + (size_t)aslot * p.a_slot_bytes));
            const int ksteps = (cb == p.cblocks - 1) ? p.last_ksteps : full_ksteps;
            int s = 0, shift_rows = p.win_shift0;
            for (int t0 = 0; t0 < taps; t0 += p.b_group) {
              const int nt = min(p.b_group, taps - t0);
              if (!ready) ptx::mbar_wait(full_bar + stage, phase);
              ptx::tc_fence_after();
              const uint32_t b_lo0 = ptx::umma_desc_lo(ptx::smem_u32(b_base + (size_t)stage * stage_bytes));
              const bool first_stage = (cb == 0 && t0 == 0);
              const int cur = stage;
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
              ready = ptx::mbar_test_wait(full_bar + stage, phase) != 0;
              if (!skip_mma) {
#define IFCB_ISSUE(KS, MS)                                                                                              \
  issue_stage<KS, MS, true>(leader, nt, first_stage, a_lo0, b_lo0, b_tap16, row16, jstride, d_tmem, (uint32_t)p.tile_n, \
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
              if (leader) ptx::umma_commit_pair(empty_bar + cur);
              __syncwarp();
            }
            if (leader) ptx::umma_commit_pair(a_empty + aslot);
            __syncwarp();
            if (++aslot == p.a_slots) { aslot = 0; aphase ^= 1; }
          }
          if (leader) ptx::umma_commit_pair(tmem_full + acc);
          __syncwarp();
          continue;
        }
        int cb = 0;
        for (int kb = 0; kb < kblocks; kb += p.b_group) {
          const int nk = min(p.b_group, kblocks - kb);
          if (!ready) ptx::mbar_wait(full_bar + stage, phase);
          ptx::tc_fence_after();
          uint32_t a_lo = ptx::umma_desc_lo(ptx::smem_u32(smem + (size_t)stage * stage_bytes));
          const int cur = stage;
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          ready = ptx::mbar_test_wait(full_bar + stage, phase) != 0;
          for (int i = 0; i < nk; ++i, a_lo += kb16) {
            const uint32_t b_lo = a_lo + (uint32_t)(a_tile_bytes >> 4);
            const int ksteps = (cb == p.cblocks - 1) ? p.last_ksteps : full_ksteps;
            const uint32_t first = (kb + i) > 0 ? 1u : 0u;
            if (!skip_mma) {
              switch (ksteps) {
                case 1: issue_tap<1, 1, true>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
                case 2: issue_tap<2, 1, true>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
                case 3: issue_tap<3, 1, true>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
                default: issue_tap<4, 1, true>(leader, a_lo, b_lo, 0u, d_tmem, 0u, desc_hi, idesc, first); break;
              }
            }
            if (++cb == p.cblocks) cb = 0;
          }
          if (leader) ptx::umma_commit_pair(empty_bar + cur);
          __syncwarp();
        }
        if (leader) ptx::umma_commit_pair(tmem_full + acc);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int quad = warp & 3;
    const int sub = (warp - 2) >> 2;
    uint8_t* stage = s_stage + (warp - 2) * kStageTile;
    const uint32_t lead_empty0 = ptx::mapa(ptx::smem_u32(tmem_empty), 0u);
    float sacc[STATS ? 3 : 1][16];
    if (STATS) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < 16; ++j) sacc[STATS ? a : 0][j] = 0.f;
    }
    int cur_n = -1;
    int local = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++local) {
      int m_tile, n_tile;
      decode_tile(p, tile, m_tiles, m_tile, n_tile);
      if (STATS && n_tile != cur_n) {
        if (cur_n >= 0) stats_flush(p, reinterpret_cast<float(&)[3][16]>(sacc), cur_n * p.tile_n, lane, sub, s_sum, s_sq);
        cur_n = n_tile;
      }
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      ptx::mbar_wait(tmem_full + acc, acc_phase);
      ptx::tc_fence_after();
      if (!(p.debug_flags & 4)) {
        const uint32_t tmem_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kAccBufCols);
        epilogue_tile<FP16, STATS>(p, tmem_acc, (m_tile * 2 + (int)rank) * tile_rows + quad * 32, n_tile * p.tile_n, lane, sub, stage,
                                   s_scale, s_shift, reinterpret_cast<float(&)[3][16]>(sacc));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(lead_empty0 + (uint32_t)(acc * 8));
    }
    if (STATS && cur_n >= 0) stats_flush(p, reinterpret_cast<float(&)[3][16]>(sacc), cur_n * p.tile_n, lane, sub, s_sum, s_sq);
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
  if (STATS) {
    for (int i = threadIdx.x; i < p.cout; i += kThreads) {
      atomicAdd(p.stats + i, (double)s_sum[i]);
      atomicAdd(p.stats + p.cout + i, (double)s_sq[i]);
    }
  }
}

constexpr int kSmemBudget = 227 * 1024;

}  // namespace

// Shared-memory plan of a layer (kp.row_bytes, tile_n, cout_pad, kh, kw, cblocks, last_ksteps set).
// Returns false if nothing fits.
bool conv_plan_smem(ConvKernelParams& kp, bool window, bool pair, int halo_rows) {
  const int row_bytes = kp.row_bytes;
  const int b_tap = kp.tile_n * row_bytes;
  const int a_tile = kBlockM * row_bytes;
  const int fixed = epilogue_smem(kp.cout_pad, kp.stats != nullptr) + 1024;
  const int taps = kp.kh * kp.kw;
  if (!window) {
    kp.m_sub = 1;
    kp.a_slots = 0;
    kp.a_slot_bytes = 0;
    kp.box_rows = kp.n_boxes = 0;
    // k-blocks per pipeline stage: two only when at least five such stages fit -- measured: prefetch
    // depth matters more than the barrier round trips saved (3 stages x 2 k-blocks lost 5-10 %)
    const int kb = a_tile + (pair ? b_tap / 2 : b_tap);
    const int kblocks = taps * kp.cblocks;
    int g = (kblocks >= 2 && (kSmemBudget - fixed) / (2 * kb) >= 5) ? 2 : 1;
    if (kp.b_group_cap > 0 && g > kp.b_group_cap) g = kp.b_group_cap;
    kp.b_group = g;
    int s = (kSmemBudget - fixed) / (g * kb);
    if (s > 10) s = 10;
    kp.stages = s;
    return s >= 2;
  }
  // WINDOW: enumerate (m accumulators per tile, A slots, taps per weight stage) and keep the
  // cheapest by a small cycle model per 128 output rows (constants measured on B200); on CTA pairs every SM holds and fills
  // half of each weight tile:
  //   MMA      taps * ksteps * max(N/2, 32 + N/4)           (smem operand port: 128 B/clk)
  //   loads    smem fill bytes / 60 B/clk                   (L2 -> SM)
  //   barriers ~250 clk of issue stall per pipeline stage   (4 MMAs of run-ahead)
  const int b_tap_w = pair ? b_tap / 2 : b_tap;          // weight-tile bytes per CTA and tap
  const int full_ks = row_bytes >> 5;
  const int ks_total = (kp.cblocks - 1) * full_ks + kp.last_ksteps;
  const int t_mma = kp.tile_n / 2 > 32 + kp.tile_n / 4 ? kp.tile_n / 2 : 32 + kp.tile_n / 4;
  const double mma_clk = (double)taps * ks_total * t_mma;
  double best = 1e30;
  bool found = false;
  for (int m = 4; m >= 1; m >>= 1) {
    if (m > kp.m_sub_cap) continue;
    if (m * kp.tile_n > kAccBufCols) continue;
    const int rows = kBlockM * m + halo_rows;
    const int n_boxes = (rows + 255) / 256;
    const int box_rows = ((rows + n_boxes - 1) / n_boxes + 7) & ~7;
    const int slot = (n_boxes * box_rows * row_bytes + 1023) & ~1023;
    for (int slots = kp.a_slots_pref; slots >= 1; --slots) {
      const int left = kSmemBudget - fixed - slots * slot;
      int gmax = 49152 / b_tap_w;
      if (gmax < 1) gmax = 1;
      if (gmax > taps) gmax = taps;
      if (kp.b_group_cap > 0 && gmax > kp.b_group_cap) gmax = kp.b_group_cap;
      for (int g = gmax; g >= 1; --g) {
        const int n_groups = (taps + g - 1) / g;
        const int gb = (taps + n_groups - 1) / n_groups;      // balanced group size
        const int b_stage = gb * b_tap_w;
        int s = left / b_stage;
        if (s > 6) s = 6;
        if (s < 2) continue;
        const double load_clk = ((double)kp.cblocks * slot + (double)kp.cblocks * taps * b_tap_w) / m / 60.0;
        const double bar_clk = (double)kp.cblocks * (n_groups + 1) * 250.0 / m;
        double score = (mma_clk > load_clk ? mma_clk : load_clk) + bar_clk;
        if (slots == 1) score += (double)kp.cblocks * slot / m / 60.0;     // A load not overlapped
        if (s == 2 && n_groups > 1) score += 0.25 * bar_clk;                // shallow weight ring
        if (score < best) {
          best = score;
          found = true;
          kp.m_sub = m;
          kp.a_slots = slots;
          kp.a_slot_bytes = slot;
          kp.box_rows = box_rows;
          kp.n_boxes = n_boxes;
          kp.stages = s;
          kp.b_group = gb;
        }
      }
    }
  }
  return found;
}

int conv_smem_bytes(const ConvKernelParams& kp, bool window, bool pair) {
  const int b_tap = kp.tile_n * kp.row_bytes;
  const int ops = window ? kp.a_slots * kp.a_slot_bytes + kp.stages * kp.b_group * (pair ? b_tap / 2 : b_tap)
                         : kp.stages * kp.b_group * (kBlockM * kp.row_bytes + (pair ? b_tap / 2 : b_tap));
  return ops + epilogue_smem(kp.cout_pad, kp.stats != nullptr) + 1024;
}

namespace {
// the opt-in shared-memory limit is a per-device function attribute: set it once per (kernel, device) to the budget
template <typename K>
int ensure_smem_attr(K kernel, bool (&done)[64]) {
  int dev = 0;
  IFCB_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !done[dev]) {
    IFCB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
  return 0;
}

template <bool WINDOW, bool FP16, bool STATS>
int launch_variant(const ConvLayer& L, const ConvKernelParams& p, int grid, int smem, cudaStream_t stream) {
  static bool done[64] = {};
  if (int rc = ensure_smem_attr(conv_umma_kernel<WINDOW, FP16, STATS>, done)) return rc;
  conv_umma_kernel<WINDOW, FP16, STATS><<<grid, kThreads, smem, stream>>>(L.tmap_a, L.tmap_b, p);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
}  // namespace

namespace {
template <bool WINDOW, bool FP16, bool STATS>
int launch_pair(const ConvLayer& L, const ConvKernelParams& p, int grid, int smem, cudaStream_t stream) {
  static bool done[64] = {};
  if (int rc = ensure_smem_attr(conv_umma_pair_kernel<WINDOW, FP16, STATS>, done)) return rc;
  // cluster dimensions (2,1,1) are compiled into the kernel (__cluster_dims__); grid is even
  conv_umma_pair_kernel<WINDOW, FP16, STATS><<<grid, kThreads, smem, stream>>>(L.tmap_a, L.tmap_b, p);
  IFCB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
}  // namespace

int launch_conv(const ConvLayer& L, int batch, cudaStream_t stream) {
  ConvKernelParams p = L.kp;
  p.rows = (long long)batch * p.rows_per_img;
  const int tile_rows = kBlockM * p.m_sub * (L.pair ? 2 : 1);
  const long long m_tiles = (p.rows + tile_rows - 1) / tile_rows;
  const long long total = m_tiles * p.n_tiles;
  if (total == 0) return 0;
  if (p.rows + tile_rows >= (1ll << 31)) {
    set_error("conv: %lld GEMM rows exceed the 32-bit row index", p.rows);
    return -1;
  }
  const int smem = conv_smem_bytes(p, L.window, L.pair);
  const bool st = p.stats != nullptr;
  if (L.pair) {
    const int pairs = sm_count() / 2;
    const int grid = 2 * (int)(total < pairs ? total : pairs);
#define IFCB_LAUNCH_PAIR(W, S) (p.fp16 ? launch_pair<W, true, S>(L, p, grid, smem, stream) : launch_pair<W, false, S>(L, p, grid, smem, stream))
    if (L.window) return st ? IFCB_LAUNCH_PAIR(true, true) : IFCB_LAUNCH_PAIR(true, false);
    return st ? IFCB_LAUNCH_PAIR(false, true) : IFCB_LAUNCH_PAIR(false, false);
#undef IFCB_LAUNCH_PAIR
  }
  const int grid = (int)(total < sm_count() ? total : sm_count());
#define IFCB_LAUNCH(W, S) (p.fp16 ? launch_variant<W, true, S>(L, p, grid, smem, stream) : launch_variant<W, false, S>(L, p, grid, smem, stream))
  if (L.window) return st ? IFCB_LAUNCH(true, true) : IFCB_LAUNCH(true, false);
  return st ? IFCB_LAUNCH(false, true) : IFCB_LAUNCH(false, false);
#undef IFCB_LAUNCH
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
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (keeps the address space known: LDS/STS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
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
