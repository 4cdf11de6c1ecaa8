// Microbenchmark (profiling aid, not product code): issue rate / throughput of
// tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16, SS operands) as a function of N, of how many
// accumulators the MMAs rotate over, and of the operand row width (128B / 64B swizzle).
// One CTA per SM; one elected lane issues `iters` MMAs back to back, then commits and waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_microbench tools/mma_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../ifcb_classifier_b200/csrc/ptx.cuh"

using namespace ifcb;

struct Cfg {
  int n;          // MMA N
  int accs;       // accumulators rotated over (1, 2, 4)
  int run;        // consecutive MMAs on the same accumulator before switching
  int row_bytes;  // 128 or 64
  int iters;      // MMAs per CTA
  int a_tiles;    // distinct A tiles rotated over (smem footprint)
  int shift;      // A start shifted by this many rows (window-style unaligned start)
  int commit_every;  // issue a tcgen05.commit (to a scratch mbarrier) after every this many MMAs (0: never)
};

__global__ void __launch_bounds__(128, 1) mma_bench(Cfg c, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t scratch_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::mbar_init(&scratch_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const bool leader = ptx::elect_one();
    const uint32_t idesc = ptx::umma_idesc_f16(128, c.n, 1);
    const uint32_t hi = ptx::umma_desc_hi(c.row_bytes);
    const uint32_t a0 = ptx::smem_u32(smem) + (uint32_t)(c.shift * c.row_bytes);
    const uint32_t a_stride = 128u * (uint32_t)c.row_bytes;              // one A tile
    const uint32_t b0 = ptx::smem_u32(smem) + 4u * 16384u + 2048u;       // B after 4 A tiles (+ shift slack)
    const int ksteps = c.row_bytes / 32;
    // 8 precomputed operand triples (the pattern repeats every 8 MMAs), issued from a tight loop
    uint32_t al[8], bl[8], dd[8];
    {
      int acc = 0, inrun = 0, at = 0, k = 0;
      for (int i = 0; i < 8; ++i) {
        al[i] = ptx::umma_desc_lo(a0 + (uint32_t)at * a_stride) + (uint32_t)(2 * k);
        bl[i] = ptx::umma_desc_lo(b0) + (uint32_t)(2 * k);
        dd[i] = tmem + (uint32_t)(acc * c.n);
        if (++k == ksteps) { k = 0; if (++at == c.a_tiles) at = 0; }
        if (++inrun == c.run) { inrun = 0; if (++acc == c.accs) acc = 0; }
      }
    }
    long long t0 = clock64();
    if (c.commit_every == 0) {
      for (int i = 0; i < c.iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (leader) ptx::umma_f16_lohi(dd[u], al[u], bl[u], hi, idesc, i > 0 ? 1u : 0u);
      }
    } else {
      for (int i = 0; i < c.iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (leader) ptx::umma_f16_lohi(dd[u], al[u], bl[u], hi, idesc, i > 0 ? 1u : 0u);
          if (((u + 1) % c.commit_every) == 0 && leader) ptx::umma_commit(&scratch_bar);
        }
      }
    }
    long long t1 = clock64();
    if (leader) ptx::umma_commit(&bar);
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (leader) {
      out_cycles[2 * blockIdx.x] = t1 - t0;
      out_cycles[2 * blockIdx.x + 1] = t2 - t0;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sizeof(long long) * 2 * 1024);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  printf("%5s %5s %4s %5s %7s %6s %6s | %10s %10s | %10s\n", "N", "accs", "run", "rowB", "a_tiles", "shift", "grid", "issue clk", "total clk",
         "floor N/2");
  const int grids[2] = {1, sms};
  for (int gi = 2; gi < 2; ++gi) {
    for (int rb = 128; rb >= 64; rb -= 64) {
      for (int n : {32, 64, 96, 128, 192, 256}) {
        for (int accs : {1, 2, 4}) {
          if (accs * n > 512) continue;
          for (int run : {1, 2, 4}) {
            if (accs == 1 && run > 1) continue;
            for (int shift : {0, 3}) {
              if (shift && (accs != 2 || run != 2)) continue;
              Cfg c{n, accs, run, rb, iters, 2, shift, 0};
              mma_bench<<<grids[gi], 128, smem>>>(c, d);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
              long long h[2 * 1024];
              cudaMemcpy(h, d, sizeof(long long) * 2 * grids[gi], cudaMemcpyDeviceToHost);
              double issue = 0, total = 0;
              for (int b = 0; b < grids[gi]; ++b) { issue += h[2 * b]; total += h[2 * b + 1]; }
              issue /= grids[gi]; total /= grids[gi];
              printf("%5d %5d %4d %5d %7d %6d %6d | %10.1f %10.1f | %10.1f\n", n, accs, run, rb, 2, shift, grids[gi], issue / iters,
                     total / iters, n / 2.0);
            }
          }
        }
      }
    }
  }
  // queue depth probe: how far can the issuing thread run ahead of execution?
  printf("\nqueue depth probe (N=256, 128B rows, 1 CTA): iters | issue clk total | exec clk total\n");
  for (int it : {8, 16, 32, 64, 128, 256}) {
    Cfg c{256, 1, 1, 128, it, 2, 0, 0};
    mma_bench<<<1, 128, smem>>>(c, d);
    cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%5d | %8lld | %8lld\n", it, h[0], h[1]);
  }
  for (int it : {8, 16, 32, 64, 128, 256}) {
    Cfg c{32, 1, 1, 128, it, 2, 0, 0};
    mma_bench<<<1, 128, smem>>>(c, d);
    cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=32 %5d | %8lld | %8lld\n", it, h[0], h[1]);
  }
  printf("\ncommit cost probe (1 CTA, 128B rows): N | commit every g MMAs | clk per MMA\n");
  for (int n : {32, 64, 128, 240}) {
    for (int g : {0, 8, 4, 2, 1}) {
      Cfg c{n, 1, 1, 128, 4096, 2, 0, g};
      mma_bench<<<1, 128, smem>>>(c, d);
      cudaDeviceSynchronize();
      long long h[2];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%5d | %3d | %8.1f\n", n, g, h[1] / 4096.0);
    }
  }
  return 0;
}
