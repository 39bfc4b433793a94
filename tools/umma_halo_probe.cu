// Probe: can one halo tile in shared memory feed all 9 taps of a 3x3 conv through UMMA descriptors
// whose start address is shifted by whole pixels? Tests SW64 / SW128 / no-swizzle K-major layouts.
// D = A_tap * I  (B = 16x16 identity), so D[m][n] = A[pixel(m)+shift][kk*16 + n].
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../conditional_score_diffusion_b200/csrc/ptx.cuh"
using namespace csd;

constexpr int TW = 8, TH = 16, HW_ = TW + 2, HH_ = TH + 2;  // halo 10 x 18
constexpr int GH = 40, GW = 24;                            // global image

struct Params {
  int variant;      // 0: SW64 (32 ch rows), 1: SW128 (64 ch rows), 2: no swizzle [kc][pix][8]
  int use_base_offset;
  int dy, dx, kk;   // tap shift in halo coords (0..2), k sub-block
  int h0, w0;
  float* out;       // [128][16]
};

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Params p) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base;               // up to 18*10*128 = 23040 B
  const uint32_t b_addr = base + 24576;       // identity 16 rows x 32 B (SW32-free: use no swizzle interleave for B)
  const uint32_t bar = base + 24576 + 1024;
  const uint32_t mbar2 = bar + 8;
  const uint32_t slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(mbar2, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(slot, 32); ptx::tmem_relinquish(); }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    uint32_t a_bytes = (p.variant == 0) ? HH_ * HW_ * 64 : (p.variant == 1 ? HH_ * HW_ * 128 : 4 * HH_ * HW_ * 16);
    ptx::mbar_arrive_expect_tx(bar, a_bytes + 16 * 32);
    if (p.variant == 2) ptx::tma_load_5d(a_addr, &mapA, bar, 0, p.w0 - 1, p.h0 - 1, 0, 0);
    else ptx::tma_load_4d(a_addr, &mapA, bar, 0, p.w0 - 1, p.h0 - 1, 0);
    ptx::tma_load_3d(b_addr, &mapB, bar, 0, 0, 0);
    ptx::mbar_wait(bar, 0);
    ptx::tcgen05_fence_after();
    uint64_t adesc;
    const int shift_rows = p.dy * HW_ + p.dx;
    if (p.variant == 0) {
      uint32_t start = a_addr + shift_rows * 64 + p.kk * 32;
      adesc = ptx::make_smem_desc(start, 16, HW_ * 64, 4);
      if (p.use_base_offset) adesc |= (uint64_t)((start >> 7) & 7) << 49;
    } else if (p.variant == 1) {
      uint32_t start = a_addr + shift_rows * 128 + p.kk * 32;
      adesc = ptx::make_smem_desc(start, 16, HW_ * 128, 2);
      if (p.use_base_offset) adesc |= (uint64_t)((start >> 7) & 7) << 49;
    } else {
      uint32_t start = a_addr + shift_rows * 16 + p.kk * 2 * (HH_ * HW_ * 16);
      adesc = ptx::make_smem_desc(start, HH_ * HW_ * 16, HW_ * 16, 0);
    }
    // B: 16 rows (n) x 16 k, K-major, no swizzle: two core matrices along N? N=16 -> 2 groups of 8 rows.
    // layout in smem from TMA 3D box {8, 16, 2}: [khalf][n][8] -> LBO = 16*16 = 256, SBO = 8*16 = 128
    uint64_t bdesc = ptx::make_smem_desc(b_addr, 256, 128, 0);
    uint32_t idesc = ptx::make_idesc_bf16_m128(16);
    ptx::mma_bf16_ss(tmem, adesc, bdesc, idesc, 0);
    ptx::mma_commit(mbar2);
    ptx::mbar_wait(mbar2, 0);
  }
  __syncthreads();
  ptx::tcgen05_fence_after();
  uint32_t r[16];
  ptx::tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16), r);
  ptx::tmem_ld_wait();
  for (int i = 0; i < 16; ++i) p.out[(warp * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 32);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn enc;

static int encode(CUtensorMap* m, int rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapSwizzle sw) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

int main() {
  cudaFree(0);
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  const int C = 64;
  // two value sets: pixel index (mod 256) and channel index
  std::vector<__nv_bfloat16> hA1(GH * GW * C), hA2(GH * GW * C), hB(2 * 16 * 8);
  for (int h = 0; h < GH; ++h) for (int w = 0; w < GW; ++w) for (int c = 0; c < C; ++c) {
    hA1[(h * GW + w) * C + c] = __float2bfloat16((float)((h * GW + w) % 251));
    hA2[(h * GW + w) * C + c] = __float2bfloat16((float)(c + 1));
  }
  // identity B stored plain [n][k] (16 x 16), TMA re-tiles it into [khalf][n][8]
  std::vector<__nv_bfloat16> hBp(16 * 16);
  for (int n = 0; n < 16; ++n) for (int k = 0; k < 16; ++k) hBp[n * 16 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA1, *dA2, *dB; float* dout;
  cudaMalloc(&dA1, hA1.size() * 2); cudaMalloc(&dA2, hA2.size() * 2); cudaMalloc(&dB, hBp.size() * 2); cudaMalloc(&dout, 128 * 16 * 4);
  cudaMemcpy(dA1, hA1.data(), hA1.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dA2, hA2.data(), hA2.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hBp.data(), hBp.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);

  CUtensorMap mapB;
  { cuuint64_t dims[3] = {8, 16, 2}; cuuint64_t str[2] = {32, 16}; cuuint32_t box[3] = {8, 16, 2};
    int r = encode(&mapB, 3, dB, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE); printf("encode B (permuted strides): %d\n", r); if (r) return 1; }

  for (int variant = 0; variant < 3; ++variant) {
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* dA = which == 0 ? dA1 : dA2;
      CUtensorMap mapA;
      int r;
      if (variant == 0) { cuuint64_t dims[4] = {(cuuint64_t)C, GW, GH, 1}; cuuint64_t str[3] = {C * 2, (cuuint64_t)GW * C * 2, (cuuint64_t)GH * GW * C * 2};
        cuuint32_t box[4] = {32, HW_, HH_, 1}; r = encode(&mapA, 4, dA, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B); }
      else if (variant == 1) { cuuint64_t dims[4] = {(cuuint64_t)C, GW, GH, 1}; cuuint64_t str[3] = {C * 2, (cuuint64_t)GW * C * 2, (cuuint64_t)GH * GW * C * 2};
        cuuint32_t box[4] = {64, HW_, HH_, 1}; r = encode(&mapA, 4, dA, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B); }
      else { cuuint64_t dims[5] = {8, GW, GH, (cuuint64_t)C / 8, 1}; cuuint64_t str[4] = {C * 2, (cuuint64_t)GW * C * 2, 16, (cuuint64_t)GH * GW * C * 2};
        cuuint32_t box[5] = {8, HW_, HH_, 4, 1}; r = encode(&mapA, 5, dA, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE); }
      if (r) { printf("variant %d: encode failed %d\n", variant, r); continue; }
      for (int ubo = 0; ubo < (variant == 2 ? 1 : 2); ++ubo)
      for (int h0 : {0, 16}) for (int dy = 0; dy < 3; ++dy) for (int dx = 0; dx < 3; ++dx) for (int kk = 0; kk < 2; ++kk) {
        Params p{variant, ubo, dy, dx, kk, h0, 8, dout};
        cudaMemset(dout, 0xff, 128 * 16 * 4);
        probe<<<1, 128, 48 * 1024>>>(mapA, mapB, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: launch error %s\n", variant, cudaGetErrorString(e)); return 1; }
        std::vector<float> h(128 * 16);
        cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
          int hl = m / TW, wl = m % TW;
          int gh = p.h0 - 1 + hl + dy, gw = p.w0 - 1 + wl + dx;
          float expv = 0.f;
          if (gh >= 0 && gh < GH && gw >= 0 && gw < GW) expv = which == 0 ? (float)((gh * GW + gw) % 251) : (float)(kk * 16 + n + 1);
          if (h[m * 16 + n] != expv) ++bad;
        }
        if (bad || (dy == 1 && dx == 1 && kk == 0 && h0 == 0))
          printf("variant %d vals %d base_off %d h0 %2d tap(%d,%d) kk %d: bad=%d  (D[0][0..3]=%g %g %g %g, D[9][0]=%g)\n", variant, which, ubo, h0, dy, dx, kk, bad,
                 h[0], h[1], h[2], h[3], h[9 * 16]);
      }
      printf("variant %d vals %d done\n", variant, which);
    }
  }
  return 0;
}
