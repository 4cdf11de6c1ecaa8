// Microbenchmark (profiling aid, not product code): cost per tcgen05.mma.cta_group::1.kind::f16 (K = 16, bf16, SWIZZLE_128B operands in
// shared memory) as a function of the operands' MAJORNESS -- K-major (the forward / data-gradient convs) vs MN-major (the weight gradient,
// whose reduction index is the pixel axis) -- of N and of M.  One CTA per SM; one elected lane issues `iters` MMAs back to back over
// 8 K steps of one operand tile pair, then commits and waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_major_microbench tools/mma_major_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../ifcb_classifier_b200/csrc/ptx.cuh"

using namespace ifcb;

struct Cfg {
  int m, n;
  int a_mn, b_mn;   // 1: MN-major
  int iters;
  int a_shift;      // A start shifted by this many 128-byte rows (WINDOW-style tap shift)
  int a_lbo;        // MN-major A: byte distance of the second 64-element block (16384 = a separate tile; 128 = the next patch pixel)
  int a_kstep;      // MN-major A: bytes per K = 16 step (2048 = 16 contiguous rows; 2304 = an 18-pixel patch row pitch)
};

__global__ void __launch_bounds__(128, 1) mma_bench(Cfg c, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
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
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) | (((uint32_t)c.n >> 3) << 17) |
                           (((uint32_t)c.m >> 4) << 24);
    // K-major: 128-byte rows (64 k per row), 8-row groups 1024 B apart; a K = 16 step advances 32 B
    // MN-major: 128-byte rows indexed by k (64 M/N elements each), 8-k-row groups SBO = 1024 B, next 64-element block LBO = 16 KB; a K = 16
    //           step advances 16 rows = 2 KB
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a0 = ptx::smem_u32(smem) + (uint32_t)c.a_shift * 128u, b0 = ptx::smem_u32(smem) + 64u * 1024u;
    const uint32_t a_lo0 = ((a0 & 0x3FFFFu) >> 4) | ((c.a_mn ? ((uint32_t)c.a_lbo >> 4) : 1u) << 16);
    const uint32_t b_lo0 = ((b0 & 0x3FFFFu) >> 4) | ((c.b_mn ? (16384u >> 4) : 1u) << 16);
    const uint32_t a_k = c.a_mn ? ((uint32_t)c.a_kstep >> 4) : 2u, b_k = c.b_mn ? 128u : 2u;
    uint32_t al[8], bl[8];
    for (int k = 0; k < 8; ++k) {
      al[k] = a_lo0 + (uint32_t)(c.a_mn ? k : (k & 3)) * a_k;        // K-major tile: 4 K steps per 128-byte row
      bl[k] = b_lo0 + (uint32_t)(c.b_mn ? k : (k & 3)) * b_k;
    }
    long long t0 = clock64();
    for (int i = 0; i < c.iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (leader) {
          asm volatile(
              "{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}\n" ::"r"(tmem),
              "r"(al[u]), "r"(bl[u]), "r"(hi), "r"(idesc), "r"(i > 0 ? 1u : 0u)
              : "memory");
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
  printf("%4s %4s %8s %8s | %10s %10s | %10s\n", "M", "N", "A", "B", "issue clk", "total clk", "math M*N/256");
  for (int m : {128, 64})
    for (int amn = 0; amn < 2; ++amn)
      for (int bmn = 0; bmn < 2; ++bmn)
        for (int n : {32, 64, 96, 128, 192, 256}) {
          Cfg c{m, n, amn, bmn, iters, 0, 16384, 2048};
          mma_bench<<<sms, 128, smem>>>(c, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s (M %d N %d A %d B %d)\n", cudaGetErrorString(e), m, n, amn, bmn); return 1; }
          static long long h[2 * 1024];
          cudaMemcpy(h, d, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
          double issue = 0, total = 0;
          for (int b = 0; b < sms; ++b) { issue += h[2 * b]; total += h[2 * b + 1]; }
          printf("%4d %4d %8s %8s | %10.1f %10.1f | %10.1f\n", m, n, amn ? "MN-major" : "K-major", bmn ? "MN-major" : "K-major", issue / sms / iters,
                 total / sms / iters, m * n / 256.0);
        }
  printf("\nMN-major A with a WINDOW-style start (row shift, LBO, K-step pitch), B MN-major, M = 128:\n%4s %6s %6s %6s | %10s\n", "N", "shift", "LBO", "kstep", "total clk");
  for (int n : {32, 64, 96})
    for (int shift : {0, 1, 3, 4})
      for (int lbo : {16384, 128, 2304})
        for (int kstep : {2048, 2304}) {
          Cfg c{128, n, 1, 1, iters, shift, lbo, kstep};
          mma_bench<<<sms, 128, smem>>>(c, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          static long long h[2 * 1024];
          cudaMemcpy(h, d, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
          double total = 0;
          for (int b = 0; b < sms; ++b) total += h[2 * b + 1];
          printf("%4d %6d %6d %6d | %10.1f\n", n, shift, lbo, kstep, total / sms / iters);
        }
  return 0;
}
