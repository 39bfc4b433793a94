// Standalone probe: 3-D fp32 TMA box loads with different inner box widths (no swizzle).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../conditional_score_diffusion_b200/csrc/ptx.cuh"
using namespace csd;

template <int BW, int BH>
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int x0, int y0, int trap_test) {
  __shared__ __align__(128) float tile[BW * BH];
  __shared__ __align__(8) unsigned long long bar;
  if (trap_test) { __trap(); }
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar), BW * BH * 4);
    ptx::tma_load_3d(ptx::smem_u32(tile), &map, ptx::smem_u32(&bar), x0, y0, 0);
  }
  const long long t0 = clock64();
  while (!ptx::mbar_try_wait(ptx::smem_u32(&bar), 0)) {
    if (clock64() - t0 > 200000000LL) { if (threadIdx.x == 0) out[0] = -12345.f; return; }
  }
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BW, int BH>
void run(EncodeFn enc, float* din, int W, int H, int P, float* dout, int x0, int y0, int trap_test = 0) {
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  cuuint32_t box[3] = {BW, BH, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, din, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("BW=%d BH=%d W=%d H=%d x0=%d y0=%d trap=%d: encode=%d ", BW, BH, W, H, x0, y0, trap_test, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return; }
  cudaMemset(dout, 0, BW * BH * 4);
  probe<BW, BH><<<1, 128>>>(map, dout, x0, y0, trap_test);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> h(BW * BH);
  cudaMemcpy(h.data(), dout, BW * BH * 4, cudaMemcpyDeviceToHost);
  // expected value at tile (r,c) = (y0+r)*W + (x0+c) if in range else 0
  int bad = 0;
  for (int rr = 0; rr < BH; ++rr) for (int c = 0; c < BW; ++c) {
    int gy = y0 + rr, gx = x0 + c;
    float exp = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? (float)(gy * W + gx) : 0.f;
    if (h[rr * BW + c] != exp) ++bad;
  }
  printf("sync=%s first=%g bad=%d\n", cudaGetErrorString(e), h[0], bad);
}

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  const int W = 160, H = 160, P = 2;
  float *din, *dout;
  std::vector<float> h(W * H * P);
  for (int i = 0; i < W * H * P; ++i) h[i] = (float)(i % (W * H));
  cudaMalloc(&din, h.size() * 4);
  cudaMalloc(&dout, 256 * 256 * 4);
  cudaMemcpy(din, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  run<64, 8>(enc, din, W, H, P, dout, 0, 0);
  run<64, 8>(enc, din, W, H, P, dout, -3, -2);
  run<68, 19>(enc, din, W, H, P, dout, 0, 0);
  run<68, 19>(enc, din, W, H, P, dout, -1, -1);
  run<132, 36>(enc, din, W, H, P, dout, 5, 5);
  run<32, 4>(enc, din, W, H, P, dout, 150, 158);
  // box larger than the tensor
  float* dsmall; cudaMalloc(&dsmall, 8 * 8 * 4 * 2);
  std::vector<float> hs(128); for (int i = 0; i < 128; ++i) hs[i] = (float)(i % 64);
  cudaMemcpy(dsmall, hs.data(), 512, cudaMemcpyHostToDevice);
  run<68, 19>(enc, dsmall, 8, 8, 2, dout, -1, -1);
  run<16, 16>(enc, dsmall, 8, 8, 2, dout, -1, -1);
  run<64, 8>(enc, din, W, H, P, dout, 0, 0, 1);  // what does __trap() report?
  return 0;
}
