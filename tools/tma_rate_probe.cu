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
  if (threadIdx.x != 0) return;
  const uint32_t s0 = (ptx::smem_u32(smem) + 1023u) & ~1023u;
  const long long t0 = clock64();
  // different CTAs walk different slabs (row block by CTA, K column by iteration) like the conv kernels do
  const int rb = (blockIdx.x * 131) % (rows_total / 128);
  for (int it = 0; it < iters + kRing; ++it) {
    const int s = it % kRing;
    if (it >= kRing) ptx::mbar_wait(ptx::smem_u32(&bars[s]), ((it / kRing) - 1) & 1);
    if (it < iters) {
      const uint32_t bar = ptx::smem_u32(&bars[s]);
      ptx::mbar_arrive_expect_tx(bar, kSlab);
      const int kc = (it * 7 + blockIdx.x) % kcols;
      if (MODE == 0) ptx::tma_load_3d(s0 + s * kSlab, &map, bar, kc * 32, rb * 128, 0);
      if (MODE == 1) ptx::tma_load_3d(s0 + s * kSlab, &map, bar, (kc / 2) * 64, rb * 128 + (kc & 1) * 64, 0);
      if (MODE == 2) ptx::bulk_load_1d(s0 + s * kSlab, base + ((long long)(rb * kcols + kc)) * kSlab, kSlab, bar);
    }
  }
  if (blockIdx.x == 0) *cyc = clock64() - t0;
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
  const char* names[3] = {"128 rows x 64 B (SW64 box)", "64 rows x 128 B (SW128 box)", "8 KB contiguous bulk copy"};
  for (int ctas : {148, 26, 1}) {
    for (int mode = 0; mode < 3; ++mode) {
      CUtensorMap map;
      cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 1};
      cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * 2 * rows};
      cuuint32_t box[3] = {mode == 1 ? 64u : 32u, mode == 1 ? 64u : 128u, 1};
      cuuint32_t es[3] = {1, 1, 1};
      enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          mode == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      for (int rep = 0; rep < 2; ++rep) {
        const size_t smem = kRing * kSlab + 1024;
        if (mode == 0) { cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<0><<<ctas, 64, smem>>>(map, d, cyc, iters, rows, kcols); }
        if (mode == 1) { cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<1><<<ctas, 64, smem>>>(map, d, cyc, iters, rows, kcols); }
        if (mode == 2) { cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<2><<<ctas, 64, smem>>>(map, d, cyc, iters, rows, kcols); }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      printf("%3d CTAs  %-30s: %.1f B/clk per SM (%lld cycles for %d slabs)\n", ctas, names[mode],
             (double)iters * kSlab / (double)*cyc, *cyc, iters);
    }
  }
  return 0;
}
