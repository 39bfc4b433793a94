// Probe 2: is the ~100-cycle per-instruction floor of small-N tcgen05.mma (tools/mma_rate_probe.cu) a dependency
// latency on the accumulator or a pipe/issue limit? Issues M=128 x N x K=16 bf16 MMAs round-robin over `nacc`
// independent TMEM accumulators (SW128 K-major operands, 4 K-steps per "stage" like the real kernels), commit every
// `per_commit` MMAs. Reports cycles per MMA against the N/2-cycle ideal.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../conditional_score_diffusion_b200/csrc/ptx.cuh"
using namespace csd;

template <int NACC>
__global__ void __launch_bounds__(192) probe(int n, int total_rounds, int per_commit, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + 65536;           // A: 4 x 16 KB blocks, B: 4 x 32 KB blocks
  const uint32_t bar = base + 65536 + 131072, bar2 = bar + 8, slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar2, 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16_m128((uint32_t)n);
    const uint32_t hi = ptx::smem_desc_hi(1024, 2);
    constexpr int acc_stride = 512 / NACC;
    const uint32_t a_lo0 = ptx::smem_desc_lo(a_addr, 16), b_lo0 = ptx::smem_desc_lo(b_addr, 16);
    long long t0 = clock64();
    int since = 0;
    uint32_t accumulate = 0;
    // one round = 4 K16 steps x NACC accumulators (the inner loops are fully unrolled: no divisions, no branches)
    for (int r = 0; r < total_rounds; ++r) {
      const uint32_t blk = (uint32_t)(r & 3);
      const uint32_t a_lo = a_lo0 + blk * (16384 >> 4), b_lo = b_lo0 + blk * (32768 >> 4);
#pragma unroll
      for (int k16 = 0; k16 < 4; ++k16) {
#pragma unroll
        for (int acc = 0; acc < NACC; ++acc)
          ptx::mma_bf16_ss(tmem + acc * acc_stride, ptx::smem_desc_join(hi, a_lo + 2 * k16),
                           ptx::smem_desc_join(hi, b_lo + 2 * k16), idesc, accumulate | (uint32_t)(k16 > 0));
      }
      accumulate = 1u;
      since += 4 * NACC;
      if (since >= per_commit) { ptx::mma_commit(bar2); since = 0; }
    }
    long long t1 = clock64();
    ptx::mma_commit(bar);
    ptx::mbar_wait(bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp >= 2) {
    ptx::mbar_wait(bar, 0);
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tcgen05_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

template <int NACC>
static void run(int n, int pc, long long* d) {
  const int total = 2048;
  cudaFuncSetAttribute(probe<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<NACC><<<148, 192, 198 * 1024>>>(n, total / (4 * NACC), pc, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d accumulators=%d commit every %2d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (ideal %d) %s\n", n, NACC, pc,
         (double)h[0] / total, (double)h[1] / total, n / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  for (int n : {64, 96, 128, 192, 256})
    for (int pc : {8, 32}) {
      run<1>(n, pc, d);
      if (n * 2 <= 512) run<2>(n, pc, d);
      if (n * 4 <= 512) run<4>(n, pc, d);
    }
  return 0;
}
