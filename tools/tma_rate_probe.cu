// GPU probe: sustained TMA fill rate per SM for 8 KB operand slabs fetched as (a) 128 rows x 64 B (SWIZZLE_64B box,
// what conv_gemm.cu issues today), (b) 64 rows x 128 B (SWIZZLE_128B box), (c) one contiguous 8 KB bulk copy.
// One CTA per SM (or fewer: argv[1] = number of CTAs), a ring of 8 buffers, one thread issuing and waiting.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate_probe tma_rate_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../conditional_score_diffusion_b200/csrc/ptx.cuh"
using namespace csd;

constexpr int kRing = 8, kSlab = 8192;

__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// MODE 0/1/2 as in the header; 3 = 256 rows x 64 B (16 KB box, half as many issues per byte);
// 4 = MODE 0 with a non-blocking test_wait spin instead of try_wait; 5 = MODE 0 issued by TWO threads (two warps,
// each with its own ring half)
template <int MODE>
__global__ void __launch_bounds__(64) probe(const __grid_constant__ CUtensorMap map, const char* base, long long* cyc,
                                            int iters, int rows_total, int kcols) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[kRing];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  constexpr int kBytes = MODE == 3 ? 2 * kSlab : kSlab;
  constexpr int kR = MODE == 3 ? kRing / 2 : (MODE == 5 ? kRing / 2 : kRing);
  if (MODE == 5) { if (threadIdx.x != 0 && threadIdx.x != 32) return; }
  else if (threadIdx.x != 0) return;
  const int who = threadIdx.x / 32;                     // MODE 5: second issuing thread uses the upper ring half
  if (MODE == 5) iters /= 2;
  const uint32_t s0 = ((ptx::smem_u32(smem) + 1023u) & ~1023u) + who * (kRing / 2) * kSlab;
  const long long t0 = clock64();
  // different CTAs walk different slabs (row block by CTA, K column by iteration) like the conv kernels do
  const int rb = (blockIdx.x * 131 + who * 7) % (rows_total / 256);
  for (int it = 0; it < iters + kR; ++it) {
    const int s = it % kR;
    const uint32_t bar = ptx::smem_u32(&bars[s + who * (kRing / 2)]);
    if (it >= kR) {
      if (MODE == 4) { while (!test_wait(bar, ((it / kR) - 1) & 1)) { } }
      else ptx::mbar_wait(bar, ((it / kR) - 1) & 1);
    }
    if (it < iters) {
      ptx::mbar_arrive_expect_tx(bar, kBytes);
      const int kc = (it * 7 + blockIdx.x) % kcols;
      if (MODE == 0 || MODE == 4 || MODE == 5) ptx::tma_load_3d(s0 + s * kSlab, &map, bar, kc * 32, rb * 128, 0);
      if (MODE == 3) ptx::tma_load_3d(s0 + s * 2 * kSlab, &map, bar, kc * 32, rb * 256, 0);
      if (MODE == 1) ptx::tma_load_3d(s0 + s * kSlab, &map, bar, (kc / 2) * 64, rb * 128 + (kc & 1) * 64, 0);
      if (MODE == 2) ptx::bulk_load_1d(s0 + s * kSlab, base + ((long long)(rb * kcols + kc)) * kSlab, kSlab, bar);
    }
  }
  if (blockIdx.x == 0 && who == 0) *cyc = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  const int rows = 4096, kcols = 54;                 // a [4096 x 1728] bf16 weight-like matrix (14 MB, L2 resident)
  const int K = kcols * 32;
  char* d;
  cudaMalloc(&d, (size_t)rows * K * 2);
  cudaMemset(d, 1, (size_t)rows * K * 2);
  long long* cyc;
  cudaMallocManaged(&cyc, 8);
  const int iters = 4000;
  const char* names[6] = {"128 rows x 64 B (SW64 box)", "64 rows x 128 B (SW128 box)", "8 KB contiguous bulk copy",
                          "256 rows x 64 B (16 KB box)", "128 x 64 B, test_wait spin", "128 x 64 B, two issuing threads"};
  for (int ctas : {148, 1}) {
    for (int mode = 0; mode < 6; ++mode) {
      CUtensorMap map;
      cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 1};
      cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * 2 * rows};
      cuuint32_t box[3] = {mode == 1 ? 64u : 32u, mode == 1 ? 64u : (mode == 3 ? 256u : 128u), 1};
      cuuint32_t es[3] = {1, 1, 1};
      enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          mode == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      for (int rep = 0; rep < 2; ++rep) {
        const size_t smem = kRing * kSlab + 1024;
        if (mode == 0) { cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<0><<<ctas, 64, smem>>>(map, d, cyc, iters, rows, kcols); }
        if (mode == 1) { cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<1><<<ctas, 64, smem>>>(map, d, cyc, iters, rows, kcols); }
        if (mode == 2) { cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<2><<<ctas, 64, smem>>>(map, d, cyc, iters, rows, kcols); }
        if (mode == 3) { cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<3><<<ctas, 64, smem>>>(map, d, cyc, iters / 2, rows, kcols); }
        if (mode == 4) { cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<4><<<ctas, 64, smem>>>(map, d, cyc, iters, rows, kcols); }
        if (mode == 5) { cudaFuncSetAttribute(probe<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<5><<<ctas, 64, smem>>>(map, d, cyc, iters, rows, kcols); }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      printf("%3d CTAs  %-30s: %.1f B/clk per SM (%lld cycles for %d slabs)\n", ctas, names[mode],
             (double)iters * kSlab / (double)*cyc, *cyc, iters);
    }
  }
  return 0;
}
