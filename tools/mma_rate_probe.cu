// Probe: issue rate of tcgen05.mma (cta_group::1, M=128, bf16, SW64 K-major operands in smem) for
// different N, with / without a commit every `per_commit` MMAs, with 1..3 CTAs per SM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../conditional_score_diffusion_b200/csrc/ptx.cuh"
using namespace csd;

__global__ void __launch_bounds__(192) probe(int n, int total_mma, int per_commit, int spin_warps, long long* out, int a_sbo, int a_shift, int b_sbo, int b_shift, int vary) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + 24576;
  const uint32_t bar = base + 24576 + 24576, bar2 = bar + 8, slot = bar + 16;
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
      const uint32_t va = vary ? (uint32_t)((i >> 1) % 9) * 64u : 0u;
      const uint64_t ad = ptx::make_smem_desc(a_addr + (i & 1) * 32 + a_shift + (a_sbo == 512 ? 0 : va), 16, a_sbo, 4);
      const uint64_t bd = ptx::make_smem_desc(b_addr + (i & 1) * 32 + b_shift + (b_sbo == 512 ? 0 : va), 16, b_sbo, 4);
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
  struct Cfg { const char* name; int n, a_sbo, a_shift, b_sbo, b_shift, vary; };
  Cfg cfgs[] = {
    {"pixels=M N=96 aligned        ", 96, 512, 0, 512, 0, 0},
    {"pixels=M N=96 A sbo640       ", 96, 640, 0, 512, 0, 0},
    {"pixels=M N=96 A sbo640 +64   ", 96, 640, 64, 512, 0, 0},
    {"pixels=M N=96 A sbo640 vary  ", 96, 640, 0, 512, 0, 1},
    {"weights=M N=256 aligned      ", 256, 512, 0, 512, 0, 0},
    {"weights=M N=256 B sbo640     ", 256, 512, 0, 640, 0, 0},
    {"weights=M N=256 B sbo640 +64 ", 256, 512, 0, 640, 64, 0},
    {"weights=M N=256 B sbo640 vary", 256, 512, 0, 640, 0, 1},
    {"weights=M N=128 B sbo640 vary", 128, 512, 0, 640, 0, 1},
  };
  for (auto& c : cfgs) for (int ctas : {1, 2}) {
    size_t smem = ctas == 1 ? 100 * 1024 : 60 * 1024;
    probe<<<148 * ctas, 192, smem>>>(c.n, total, 2, 4, d, c.a_sbo, c.a_shift, c.b_sbo, c.b_shift, c.vary);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%s ctas/SM=%d: %6.1f cyc/MMA (ideal %d) %s\n", c.name, ctas, (double)h[1] / total, 128 * c.n / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
