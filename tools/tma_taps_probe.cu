// Probe: ONE TMA op delivering three weight slabs (taps t, t+1, t+2 of one 32-channel chunk) through a 3-D tensor map
// whose third dimension (tap) has a SMALLER stride than the second (row): Wt[row][tap * kstep + k] viewed as
// (k, row, tap) with strides (pitch, kstep * 2 bytes). Does cuTensorMapEncodeTiled accept it and does the box land as
// [tap][row][64 B] - three consecutive 8 KB slabs in the layout the UMMA descriptors expect?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../conditional_score_diffusion_b200/csrc/ptx.cuh"
using namespace csd;

__global__ void probe(const __grid_constant__ CUtensorMap map, unsigned short* out, int k0, int r0, int t0) {
  extern __shared__ __align__(1024) unsigned char raw[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar), 3 * 128 * 64);
    ptx::tma_load_3d(base, &map, ptx::smem_u32(&bar), k0, r0, t0);
  }
  const long long c0 = clock64();
  while (!ptx::mbar_try_wait(ptx::smem_u32(&bar), 0)) {
    if (clock64() - c0 > 200000000LL) { if (threadIdx.x == 0) out[0] = 0xDEAD; return; }
  }
  const unsigned short* tile = reinterpret_cast<const unsigned short*>(raw + (base - ptx::smem_u32(raw)));
  for (int i = threadIdx.x; i < 3 * 128 * 32; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&enc), cudaEnableDefault, &q);
  const int rows = 128, kstep = 96, taps = 9, K = taps * kstep;
  std::vector<unsigned short> h((size_t)rows * K);
  for (int r = 0; r < rows; ++r) for (int c = 0; c < K; ++c) h[(size_t)r * K + c] = (unsigned short)((r * 131 + c * 7) & 0xFFFF);
  unsigned short *d, *dout;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&dout, 3 * 128 * 32 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  for (int sw = 0; sw < 2; ++sw) {
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)kstep, (cuuint64_t)rows, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)kstep * 2};
    cuuint32_t box[3] = {32, 128, 3};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sw ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle=%s: encode=%d\n", sw ? "64B" : "none", (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int k0 = 32, r0 = 0, t0 = 3;
    cudaMemset(dout, 0, 3 * 128 * 32 * 2);
    probe<<<1, 128, 40 * 1024>>>(map, dout, k0, r0, t0);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<unsigned short> o(3 * 128 * 32);
    cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < 3; ++t) for (int rr = 0; rr < 128; ++rr) for (int j = 0; j < 32; ++j) {
      // SWIZZLE_64B: 16-byte unit u of row rr sits at unit u ^ ((rr >> 1) & 3)
      const int u = j / 8, w = j % 8;
      const int pu = sw ? (u ^ ((rr >> 1) & 3)) : u;
      const unsigned short got = o[(size_t)t * 128 * 32 + rr * 32 + pu * 8 + w];
      const unsigned short exp = h[(size_t)(r0 + rr) * K + (t0 + t) * kstep + k0 + j];
      if (got != exp) ++bad;
    }
    printf("  sync=%s first=%u bad=%d of %d\n", cudaGetErrorString(e), (unsigned)o[0], bad, 3 * 128 * 32);
  }
  return 0;
}
