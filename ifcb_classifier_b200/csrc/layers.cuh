// Internal layer records of a network plan (host side) and the kernel parameter
// blocks they launch with.  Not part of the public ABI (include/ifcb_b200.h).
#pragma once
#include <cuda.h>
#include <vector>
#include "common.cuh"
#include "../../include/ifcb_b200.h"

namespace ifcb {

struct ConvKernelParams {
  int M;                       // batch * P * Q (set per launch)
  int PQ, Q;                   // output pixels per image, output width
  int kh, kw, stride_h, stride_w, pad_h, pad_w;
  int cblocks;                 // ceil(Cin / 64)
  int last_ksteps;             // K=16 MMA steps in the last channel block (1..4)
  int tile_n, n_tiles, stages;
  int cout_pad;                // n_tiles * tile_n
  int fp16;                    // 16-bit operand/activation format: 0 bf16, 1 fp16
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  int res_ld;
  int n_seg;
  int seg_begin[IFCB_MAX_SEGMENTS], seg_end[IFCB_MAX_SEGMENTS], seg_ld[IFCB_MAX_SEGMENTS],
      seg_relu[IFCB_MAX_SEGMENTS];
  __nv_bfloat16* seg_out[IFCB_MAX_SEGMENTS];
};

struct ConvLayer {
  CUtensorMap tmap_a;          // im2col map over the NHWC input view
  CUtensorMap tmap_b;          // tiled map over the packed weights
  ConvKernelParams kp;
  int batch_cap;
};

struct StemLayer {
  ifcb_stem_desc d;
  int P, Q;
};

struct PoolLayer {
  ifcb_pool_desc d;
  int P, Q;
};

struct HeadLayer {
  ifcb_head_desc d;
};

int launch_conv(const ConvLayer& L, int batch, cudaStream_t stream);
int conv_pick_stages(int tile_n, int cout_pad);
int launch_im2col_probe(const ConvLayer& L, int c, int w, int h, int n, int off_w, int off_h, void* d_out,
                        cudaStream_t stream);
int conv_smem_bytes(int tile_n, int stages, int cout_pad);
int launch_stem(const StemLayer& L, int batch, cudaStream_t stream);
int launch_pool(const PoolLayer& L, int batch, cudaStream_t stream);
int launch_head(const HeadLayer& L, int batch, cudaStream_t stream);

}  // namespace ifcb
