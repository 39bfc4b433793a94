// Probe: what does a kernel -> kernel dependency cost inside a CUDA graph on this GPU, with and without programmatic
// dependent launch (PDL)? A chain of N small dependent kernels (each reads what the previous one wrote) is captured into
// a graph and replayed; the per-node time is the launch-to-launch floor that every one of the ~570 launches of a PC step
// pays on top of its own work. Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/pdl_probe tools/pdl_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <bool kWait, bool kTrigger>
__global__ void node_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int spin) {
  if (kTrigger) asm volatile("griddepcontrol.launch_dependents;");
  // prologue work that does not depend on the previous kernel (stands for barrier init / TMEM alloc / descriptor prefetch)
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if (kWait) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] + 1.f;
}

template <bool kWait, bool kTrigger>
static float run_chain(int nodes, int blocks, int threads, int spin, bool pdl, float* a, float* b, int n, int reps) {
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int k = 0; k < nodes; ++k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(threads);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    const float* in = (k & 1) ? b : a;
    float* out = (k & 1) ? a : b;
    cudaError_t e = cudaLaunchKernelEx(&cfg, node_kernel<kWait, kTrigger>, in, out, n, spin);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  if (cudaStreamEndCapture(s, &g) != cudaSuccess) { printf("capture failed\n"); exit(1); }
  if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { printf("instantiate failed\n"); exit(1); }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int r = 0; r < 3; ++r) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEventRecord(e0, s);
  for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  cudaStreamDestroy(s);
  return ms * 1e3f / (reps * nodes);   // us per node
}

int main() {
  const int n = 148 * 256;
  float *a, *b;
  cudaMalloc(&a, n * sizeof(float));
  cudaMalloc(&b, n * sizeof(float));
  cudaMemset(a, 0, n * sizeof(float));
  const int nodes = 570, reps = 20;
  for (int spin : {0, 2000, 6000}) {
    for (int blocks : {148, 592}) {
      const float plain = run_chain<false, false>(nodes, blocks, 256, spin, false, a, b, n, reps);
      const float pdl_wait = run_chain<true, false>(nodes, blocks, 256, spin, true, a, b, n, reps);
      const float pdl_trig = run_chain<true, true>(nodes, blocks, 256, spin, true, a, b, n, reps);
      printf("prologue %5d cycles, %3d CTAs: plain %.2f us/node | PDL (wait only) %.2f | PDL (early trigger + wait) %.2f\n",
             spin, blocks, plain, pdl_wait, pdl_trig);
    }
  }
  // correctness of the chain under PDL: 570 nodes x (3 + 20) launches of +1 each, per variant - just check it is finite
  float h = 0.f;
  cudaMemcpy(&h, (nodes & 1) ? b : a, sizeof(float), cudaMemcpyDeviceToHost);
  printf("value after all chains: %.0f (err %s)\n", h, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
