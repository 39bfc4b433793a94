// GPU probe: issue rate of MUFU.TANH / MUFU.EX2 / MUFU.RCP per SM sub-partition (one warp per SMSP, 8 independent
// chains per thread). Prints cycles per warp-instruction.   nvcc -arch=sm_100a -O3 -o mufu_probe mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void probe(float* out, long long* cyc, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = 0.1f * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  const int iters = 4096;
  const char* names[4] = {"tanh.approx", "ex2.approx", "rcp.approx", "fma"};
  for (int warps = 1; warps <= 4; warps *= 2)
    for (int op = 0; op < 4; ++op) {
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) probe<0><<<148, 128 * warps>>>(out, cyc, iters);
        if (op == 1) probe<1><<<148, 128 * warps>>>(out, cyc, iters);
        if (op == 2) probe<2><<<148, 128 * warps>>>(out, cyc, iters);
        if (op == 3) probe<3><<<148, 128 * warps>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      printf("%d warp(s)/SMSP %-12s: %.2f cycles per warp-instruction per SMSP\n", warps, names[op],
             (double)*cyc / (iters * 8.0 * warps));
    }
  return 0;
}
