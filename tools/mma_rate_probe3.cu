// Probe 3: tcgen05.mma M=128 x N x K16 bf16 issue rate with the TRANSPOSED KERNEL's operand geometry: SWIZZLE_64B
// K-major rows of 64 bytes (2 K-steps per row block), A = 128 weight rows (SBO 512), B = N pixel rows whose 8-row
// groups are `sbo_b` bytes apart (512 = dense, 640 = the 10-pixel halo pitch). Compared with probe 2 (SWIZZLE_128B):
// is the 170-190 cycles per N=256 MMA the kernel measures a property of 64-byte-row operands?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../conditional_score_diffusion_b200/csrc/ptx.cuh"
using namespace csd;

template <int KSTEPS>
__global__ void __launch_bounds__(192) probe(int n, int total_rounds, int sbo_b, int layout, int per_commit, long long* out,
                                             int b_shift_rows = 0, int rotate_shift = 0, int random_data = 0) {
  constexpr int row_bytes = KSTEPS * 32;
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + 65536;           // A: 4 x 16 KB blocks, B: 4 x 32 KB blocks
  const uint32_t bar = base + 65536 + 131072, bar2 = bar + 8, slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (random_data) {
    // operands of realistic magnitude instead of whatever the shared memory held (zeros on a fresh context): the
    // tensor pipe's sustained rate under a power cap depends on the data it multiplies
    uint32_t* w = reinterpret_cast<uint32_t*>(raw + (base - ptx::smem_u32(raw)));
    uint32_t st = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
    for (int i = threadIdx.x; i < (65536 + 131072) / 4; i += blockDim.x) {
      st = st * 1664525u + 1013904223u;
      const uint32_t lo = 0x3C00u + ((st >> 9) & 0x3FFu) + ((st >> 3) & 0x8000u);      // bf16 in [2^-7, 2^-5), random sign
      st = st * 1664525u + 1013904223u;
      const uint32_t hi = 0x3C00u + ((st >> 9) & 0x3FFu) + ((st >> 3) & 0x8000u);
      w[i] = lo | (hi << 16);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar2, 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16_m128((uint32_t)n);
    const uint32_t a_hi = ptx::smem_desc_hi(8 * row_bytes, layout);
    const uint32_t b_hi = ptx::smem_desc_hi(sbo_b, layout);
    const uint32_t a_lo0 = ptx::smem_desc_lo(a_addr, 16), b_lo0 = ptx::smem_desc_lo(b_addr, 16);
    long long t0 = clock64();
    uint32_t accumulate = 0;
    int since = 0;
    for (int r = 0; r < total_rounds; ++r) {
      const uint32_t blk = (uint32_t)(r & 3);
      // b_shift_rows: start the B operand that many 64-byte rows into its block (the transposed kernel's tap shifts:
      // (ky * 10 + kx) rows, not a multiple of the 8-row swizzle atom); rotate_shift: cycle through the 9 tap shifts
      const int tap = rotate_shift ? (r % 9) : 0;
      const uint32_t shift = rotate_shift ? (uint32_t)((tap / 3) * 10 + tap % 3) : (uint32_t)b_shift_rows;
      const uint32_t a_lo = a_lo0 + blk * (16384 >> 4), b_lo = b_lo0 + blk * (32768 >> 4) + ((shift * row_bytes) >> 4);
#pragma unroll
      for (int k16 = 0; k16 < KSTEPS; ++k16)
        ptx::mma_bf16_ss(tmem, ptx::smem_desc_join(a_hi, a_lo + 2 * k16), ptx::smem_desc_join(b_hi, b_lo + 2 * k16), idesc,
                         accumulate | (uint32_t)(k16 > 0));
      accumulate = 1u;
      since += KSTEPS;
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

static int g_mmas = 2048;
static int g_shift = 0, g_rotate = 0, g_random = 0;
template <int KSTEPS>
static void run(int n, int sbo_b, int layout, int pc, const char* name, long long* d) {
  const int mmas = g_mmas;
  cudaFuncSetAttribute(probe<KSTEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<KSTEPS><<<148, 192, 198 * 1024>>>(n, mmas / KSTEPS, sbo_b, layout, pc, d, g_shift, g_rotate, g_random);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%s N=%3d SBO_B=%4d commit every %2d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (ideal %d) %s\n", name, n, sbo_b, pc,
         (double)h[0] / mmas, (double)h[1] / mmas, n / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char** argv) {
  const int sw64 = argc > 1 ? atoi(argv[1]) : 4, sw128 = argc > 2 ? atoi(argv[2]) : 2;
  long long* d; cudaMalloc(&d, 16);
  if (argc > 5) {
    // data mode (argv[5] = anything): zeros / leftovers vs random bf16 operands, long runs (power limiter active)
    g_mmas = 400000;
    for (int rnd : {0, 1, 0, 1})
      for (int n : {160, 256}) {
        g_random = rnd;
        printf("%s operands: ", rnd ? "random  " : "leftover");
        run<2>(n, 640, sw64, 2, "SW64", d);
      }
    return 0;
  }
  if (argc > 4) {
    // operand-shift mode (argv[4] = anything): B operand started 0..12 rows of 64 bytes into its block, then the 9 tap
    // shifts of the transposed kernel in rotation
    g_mmas = 8192;
    for (int n : {160, 256}) {
      for (int sh : {0, 1, 2, 4, 8, 10, 11, 12}) {
        g_shift = sh; g_rotate = 0;
        printf("shift %2d rows: ", sh);
        run<2>(n, 640, sw64, 2, "SW64", d);
      }
      g_rotate = 1;
      printf("9 tap shifts : ");
      run<2>(n, 640, sw64, 2, "SW64", d);
      g_rotate = 0;
    }
    return 0;
  }
  if (argc > 3) {
    // sustained mode: argv[3] MMAs per launch (e.g. 400000 = tens of milliseconds of tensor work on every SM): does the
    // rate per instruction change once the power limiter has had time to act?
    g_mmas = atoi(argv[3]);
    for (int rep = 0; rep < 3; ++rep)
      for (int n : {96, 160, 224, 256}) run<2>(n, 640, sw64, 2, "SW64 sustained", d);
    return 0;
  }
  for (int n : {96, 160, 224, 256})
    for (int pc : {2, 4, 8, 16}) {
      run<2>(n, 512, sw64, pc, "SW64 ", d);
      run<2>(n, 640, sw64, pc, "SW64 ", d);      // 256 rows: 32 groups x 640 B = 20 KB < the 32 KB block
      if (pc >= 4) run<4>(n, 1024, sw128, pc, "SW128", d);
    }
  return 0;
}
