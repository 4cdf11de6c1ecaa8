// Internal layer records of a network plan (host side) and the kernel parameter
// blocks they launch with.  Not part of the public ABI (include/ifcb_b200.h).
#pragma once
#include <cuda.h>
#include <vector>
#include "common.cuh"
#include "../../include/ifcb_b200.h"

namespace ifcb {

struct ConvKernelParams {
  long long rows;              // GEMM rows of this launch (set per launch): batch * rows_per_img
  int rows_per_img, row_w;     // IM2COL: P*Q, Q        WINDOW: Hp*Wp, Wp (padded input)
  unsigned long long magic_img, magic_w;   // floor((2^64-1)/d)+1 for d = rows_per_img, row_w (0: d == 1)
  int identity_rows;           // GEMM row i is output pixel i of every (unpadded) destination: skip the row mapping
  int P, Q;                    // valid output extent per image
  int kh, kw, stride_h, stride_w, pad_h, pad_w;
  int row_bytes;               // operand row = one swizzle span: 128 (64 channels) or 64 (32 channels, Cin <= 32)
  int cblocks;                 // ceil(Cin / channels per row)
  int last_ksteps;             // K=16 MMA steps in the last channel block (1..4)
  int tile_n, n_tiles, stages;
  int cout_pad;                // n_tiles * tile_n
  int m_sub;                   // 128-row accumulators per tile (WINDOW: 1, 2 or 4)
  int a_slots, a_slot_bytes, box_rows, n_boxes;   // WINDOW: A patch ring
  int b_group;                 // WINDOW: filter taps per weight pipeline stage
  int win_shift0;              // WINDOW: (in_pad_h-pad_h)*Wp + (in_pad_w-pad_w) rows
  int m_sub_cap, a_slots_pref, b_group_cap;   // tuning knobs (env IFCB_CONV_MSUB / _ASLOTS / _BGROUP)
  int debug_flags;             // profiling only (IFCB_CONV_DEBUG): 1 skip TMA loads, 2 skip MMAs, 4 skip epilogue
  int fp16;                    // 16-bit operand/activation format: 0 bf16, 1 fp16
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  int res_ld, res_pad_h, res_pad_w;
  int n_seg;
  int seg_begin[IFCB_MAX_SEGMENTS], seg_end[IFCB_MAX_SEGMENTS], seg_ld[IFCB_MAX_SEGMENTS],
      seg_relu[IFCB_MAX_SEGMENTS], seg_pad_h[IFCB_MAX_SEGMENTS], seg_pad_w[IFCB_MAX_SEGMENTS];
  __nv_bfloat16* seg_out[IFCB_MAX_SEGMENTS];
  double* stats;               // TRAIN: float64 [2][cout] accumulators of sum / sum of squares per output channel (NULL: off)
  int cout;                    // real output channels (stats columns)
  int n_major;                 // tile order: 0 = pixel tiles outermost (m, n), 1 = channel tiles outermost (stats: a CTA keeps its columns)
};

struct ConvLayer {
  CUtensorMap tmap_a;          // IM2COL: im2col map over the NHWC input; WINDOW: tiled map over [N*Hp*Wp, C]
  CUtensorMap tmap_b;          // tiled map over the packed weights
  ConvKernelParams kp;
  int batch_cap;
  bool window;
  bool pair;                   // IM2COL on CTA pairs (cta_group::2): tmap_b box holds tile_n/2 rows
};

struct StemLayer {
  ifcb_stem_desc d;
  int P, Q;
  std::vector<float> h_const;  // gray 3x3 fast path: host copy of [9*Cout] weights, [Cout] scale, [Cout] shift (kernel parameter)
};

struct PoolLayer {
  ifcb_pool_desc d;
  int P, Q;
};

struct HeadLayer {
  ifcb_head_desc d;
};

// cuTensorMapEncode* entry points (resolved through the runtime: no link-time libcuda dependency)
#include <cudaTypedefs.h>
extern PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled;
extern PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col;
int resolve_driver();

int launch_conv(const ConvLayer& L, int batch, cudaStream_t stream);
bool conv_plan_smem(ConvKernelParams& kp, bool window, bool pair, int halo_rows);
int launch_im2col_probe(const ConvLayer& L, int c, int w, int h, int n, int off_w, int off_h, void* d_out,
                        cudaStream_t stream);
int conv_smem_bytes(const ConvKernelParams& kp, bool window, bool pair);
int launch_stem(const StemLayer& L, int batch, cudaStream_t stream);
int launch_pool(const PoolLayer& L, int batch, cudaStream_t stream);
int launch_head(const HeadLayer& L, int batch, int out_row, cudaStream_t stream);

}  // namespace ifcb
