// Probe: issue rate of tcgen05.mma (cta_group::1, M=128, bf16, SW64 K-major operands in smem) for
// different N, with / without a commit every `per_commit` MMAs, with 1..3 CTAs per SM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../conditional_score_diffusion_b200/csrc/ptx.cuh"
using namespace csd;

__global__ void __launch_bounds__(192) probe(int n, int total_mma, int per_commit, int spin_warps, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + 16384;
  const uint32_t bar = base + 16384 + 32768, bar2 = bar + 8, slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar2, 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(slot, 256); ptx::tmem_relinquish(); }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16_m128((uint32_t)n);
    long long t0 = clock64();
    int since = 0;
    for (int i = 0; i < total_mma; ++i) {
      const uint64_t ad = ptx::make_smem_desc(a_addr + (i & 1) * 32 + ((i >> 1) & 1) * 8192, 16, 512, 4);
      const uint64_t bd = ptx::make_smem_desc(b_addr + (i & 1) * 32, 16, 512, 4);
      ptx::mma_bf16_ss(tmem, ad, bd, idesc, i > 0);
      if (per_commit > 0 && ++since == per_commit) { ptx::mma_commit(bar2); since = 0; }
    }
    long long t1 = clock64();
    ptx::mma_commit(bar);
    ptx::mbar_wait(bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp >= 2 && warp < 2 + spin_warps) {
    ptx::mbar_wait(bar, 0);   // epilogue-style waiters
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tcgen05_fence_after(); ptx::tmem_dealloc(tmem, 256); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int total = 512;
  for (int ctas_per_sm : {1, 2}) for (int n : {96, 192, 256}) for (int pc : {0, 2, 8}) for (int spin : {0, 4}) {
    size_t smem = ctas_per_sm == 1 ? 100 * 1024 : 60 * 1024;
    probe<<<148 * ctas_per_sm, 192, smem>>>(n, total, pc, spin, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("ctas/SM=%d N=%3d commit_every=%d spin_warps=%d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (ideal %d) %s\n", ctas_per_sm, n, pc, spin,
           (double)h[0] / total, (double)h[1] / total, 128 * n / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
