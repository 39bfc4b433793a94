// Weight gradients of the convolutions / 1x1 projections of the score network (training backward).
//
// Reference: the reference has no hand-written backward - `loss.backward()` (losses.py:345-407 via Lightning)
// reaches cuDNN's wgrad for nn.Conv2d (models/layers.py:100-132) and ATen's einsum backward for NIN
// (models/layers.py:546-564). The contraction is
//     dW[co, ci, ky, kx] = sum over (image, output pixel o) of  g[o, co] * a[o*stride + k - pad, ci]
// i.e. a GEMM whose K dimension is the PIXEL index. tcgen05 takes K-major operands from shared memory, so
// both tensors are first re-laid-out from NHWC into a "pixel-major" form (one row per channel, pixels
// contiguous, on a zero-padded (H+2) x Wp grid, Wp = ceil8(W+2)):
//   - the activation is written in three copies shifted by kx-1 = -1, 0, +1 pixels, so that every tap's
//     operand starts on a 16-byte boundary (TMA box rows must be 16-byte aligned in global memory; a ky
//     shift is a whole padded row = a multiple of 8 elements);
//   - the output gradient is written once, on the INPUT grid at (o*stride + 1 - pad): for the stride-2
//     convolutions this is the zero-stuffed gradient, which turns their wgrad into the stride-1 form.
// Because g is zero on every padding position of the grid, products that involve a wrapped / out-of-image
// activation position vanish, and the flat shift (ky-1)*Wp selects the tap.
//
// wgrad_gemm_kernel: one CTA = (128 output channels) x (n_tile input channels) x (one tap) x (one split of the
// batch); K loop over 64-pixel chunks (128-byte swizzled rows) through a TMA ring; fp32 accumulators in TMEM;
// partial sums per split are written to a scratch tensor and reduced in a fixed order by wgrad_reduce_kernel
// (deterministic: no atomics), which also applies the layer's output scale and scatters into the reference's
// parameter layout ([Cout, Cin, kh, kw] for nn.Conv2d, [in, out] for NIN.W).
#include "common.cuh"
#include "ptx.cuh"
#include "tensormap.cuh"
#include "../../include/csd_b200.h"

#include <algorithm>
#include <cstdlib>

namespace csd {

// One 16-column chunk of an accumulator row to the fp32 partial buffer: 4 x 16-byte stores when the row segment is
// complete and aligned, statically indexed predicated stores otherwise (a dynamically indexed r[] would be placed
// in local memory and every chunk would round-trip through the stack).
__device__ __forceinline__ void store_partial_chunk(float* dst, const uint32_t (&r)[16], int cnt, bool zero) {
  if (cnt == 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      reinterpret_cast<float4*>(dst)[i] =
          zero ? make_float4(0.f, 0.f, 0.f, 0.f)
               : make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                             __uint_as_float(r[4 * i + 3]));
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < cnt) dst[i] = zero ? 0.f : __uint_as_float(r[i]);
  }
}


// ---------------------------------------------------------------------------------------------------------
// NHWC bf16 -> pixel-major copies
// ---------------------------------------------------------------------------------------------------------
struct PixMajorParams {
  const __nv_bfloat16* src;
  int pitch, c_off, c_cnt;
  int batch, h, w;             // source extent
  int stride, offset;          // grid position of source pixel (y, x) = (y*stride + offset, x*stride + offset)
  int wp;                      // padded grid row width
  long long q;                 // padded grid pixels per image = (grid_h + 2) * wp
  int ips;                     // images per split
  int margin;                  // first data column of a row
  long long row_pitch;         // elements per channel row = margin + kp + margin
  int ncopies;                 // 1 (centre only) or 3 (kx = 0, 1, 2)
  long long copy_stride;       // elements between copies
  __nv_bfloat16* out;          // [copies][splits][c_cnt][row_pitch]
};

// grid (ceil(c_cnt / 64), h, batch), block 256. One source row (b, y) x 64 channels per CTA: 64-pixel x 64-channel
// tiles are read along channels (16-byte vectors), transposed through shared memory, and written along pixels.
__global__ void __launch_bounds__(256) pixmajor_kernel(const PixMajorParams p) {
  __shared__ __nv_bfloat16 tile[64][66];   // [channel][pixel], +2 padding: conflict-free column reads
  const int c0 = blockIdx.x * 64;
  const int y = blockIdx.y, b = blockIdx.z;
  const int split = b / p.ips, bl = b % p.ips;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gy = y * p.stride + p.offset;
  const long long row0 = (long long)p.margin + (long long)bl * p.q + (long long)(gy + 1) * p.wp;
  for (int x0 = 0; x0 < p.w; x0 += 64) {
    // load: thread -> (pixel = t / 8 (+32), 8-channel vector = t % 8)
    const int cv = threadIdx.x & 7;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int px = (threadIdx.x >> 3) + half * 32;
      const int x = x0 + px;
      const int c = c0 + cv * 8;
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
      if (x < p.w && c < p.c_cnt) {
        const __nv_bfloat16* sp = p.src + (((long long)b * p.h + y) * p.w + x) * p.pitch + p.c_off + c;
        if (c + 8 <= p.c_cnt && ((p.c_off | p.pitch) & 7) == 0) {
          bf16x8 v = *reinterpret_cast<const bf16x8*>(sp);
          unpack8(v, f);
        } else {
          for (int i = 0; i < 8 && c + i < p.c_cnt; ++i) f[i] = __bfloat162float(sp[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) tile[cv * 8 + i][px] = __float2bfloat16_rn(f[i]);
    }
    __syncthreads();
    // store: warp -> 8 channels; lane -> pixels lane and lane + 32
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int cl = warp * 8 + k;
      const int c = c0 + cl;
      if (c >= p.c_cnt) break;
      __nv_bfloat16* rowp = p.out + ((long long)split * p.c_cnt + c) * p.row_pitch + row0;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int px = lane + half * 32;
        const int x = x0 + px;
        if (x >= p.w) continue;
        const __nv_bfloat16 v = tile[cl][px];
        const int gx = x * p.stride + p.offset;
        if (p.ncopies == 1) {
          rowp[gx + 1] = v;
        } else {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) rowp[(long long)kx * p.copy_stride + gx + 2 - kx] = v;
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// split-K wgrad GEMM on tcgen05
// ---------------------------------------------------------------------------------------------------------
constexpr int kWK = 64;                       // pixels (K elements) per stage: 128-byte rows, SWIZZLE_128B
constexpr int kWM = 128;                      // UMMA M (output channels per CTA)
constexpr int kWABytes = kWM * kWK * 2;       // 16 KB
constexpr int kWThreads = 224;                // warp 0: g producer, warp 1: MMA, warps 2-5: epilogue, warp 6: a producer
constexpr int kWMaxStages = 8;
constexpr uint32_t kLayoutSw128 = 2;

struct WgradParams {
  int taps, wp, margin, k_chunks;
  int cout, cin, n_tile, n_tiles;
  int num_stages, tmem_cols;
  uint32_t stage_bytes, b_bytes;
  float* partial;   // [splits][taps][cout][cin]
};

__global__ void __launch_bounds__(kWThreads, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapA0,
                  const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
                  const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + p.num_stages * p.stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWMaxStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kWMaxStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kWMaxStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x / p.n_tiles, nt = blockIdx.x % p.n_tiles;
  const int m0 = mt * kWM, n0 = nt * p.n_tile;
  const int tap = blockIdx.y, split = blockIdx.z;
  const int ky = p.taps == 9 ? tap / 3 : 1, kx = p.taps == 9 ? tap % 3 : 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapG);
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(full_bar(s), 2);    // one arrive.expect_tx per producer
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {                       // ===== producer of the output-gradient operand (M side) =====
      uint32_t stage = 0, par = 1;
      for (int k = 0; k < p.k_chunks; ++k) {
        ptx::mbar_wait(empty_bar(stage), par);
        ptx::mbar_arrive_expect_tx(full_bar(stage), (uint32_t)kWABytes);
        ptx::tma_load_3d(smem_base + stage * p.stage_bytes, &mapG, full_bar(stage), p.margin + k * kWK, m0, split);
        if (++stage == (uint32_t)p.num_stages) { stage = 0; par ^= 1u; }
      }
    }
  } else if (warp == 6) {
    if (lane == 0) {                       // ===== producer of the activation operand (N side) =====
      const CUtensorMap* mapA = kx == 0 ? &mapA0 : (kx == 1 ? &mapA1 : &mapA2);
      ptx::prefetch_tensormap(mapA);
      const int col0 = p.margin + (ky - 1) * p.wp;
      uint32_t stage = 0, par = 1;
      for (int k = 0; k < p.k_chunks; ++k) {
        ptx::mbar_wait(empty_bar(stage), par);
        ptx::mbar_arrive_expect_tx(full_bar(stage), p.b_bytes);
        ptx::tma_load_3d(smem_base + stage * p.stage_bytes + kWABytes, mapA, full_bar(stage), col0 + k * kWK, n0, split);
        if (++stage == (uint32_t)p.num_stages) { stage = 0; par ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                       // ===== MMA issuer =====
      const uint32_t idesc = ptx::make_idesc_bf16_m128((uint32_t)p.n_tile);
      const uint32_t hi = ptx::smem_desc_hi(1024, kLayoutSw128);
      uint32_t stage = 0, par = 0, accumulate = 0;
      for (int k = 0; k < p.k_chunks; ++k) {
        ptx::mbar_wait(full_bar(stage), par);
        ptx::tcgen05_fence_after();
        const uint32_t a_lo = ptx::smem_desc_lo(smem_base + stage * p.stage_bytes, 16);
        const uint32_t b_lo = a_lo + (kWABytes >> 4);
#pragma unroll
        for (int kk = 0; kk < kWK / 16; ++kk) {   // 32 bytes of K per instruction inside the 128-byte swizzled row
          ptx::mma_bf16_ss(tmem_base, ptx::smem_desc_join(hi, a_lo + 2 * kk), ptx::smem_desc_join(hi, b_lo + 2 * kk), idesc,
                           accumulate);
          accumulate = 1u;
        }
        ptx::mma_commit(empty_bar(stage));
        if (++stage == (uint32_t)p.num_stages) { stage = 0; par ^= 1u; }
      }
      ptx::mma_commit(tmem_full_bar);
    }
  } else {
    // ===== epilogue (warps 2..5): TMEM lane = output channel =====
    const int q = warp & 3;
    const int co = m0 + q * 32 + lane;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tcgen05_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    float* orow = p.partial + (((long long)split * p.taps + tap) * p.cout + co) * p.cin;
    const int ncols = min(p.n_tile, p.cin - n0);
    for (int col = 0; col < ncols; col += 16) {
      uint32_t r[16];
      __syncwarp();
      ptx::tmem_ld_x16(t_row + col, r);
      ptx::tmem_ld_wait();
      if (co < p.cout) store_partial_chunk(orow + n0 + col, r, min(16, ncols - col), false);
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// dW[co*s_co + (ci_off+ci)*s_ci + tap*s_tap] (+)= scale * sum_split partial[split][tap][co][ci]
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int taps, int cout, int cin, float scale,
                    float* __restrict__ dw, long long s_co, long long s_ci, long long s_tap, int ci_off, int accumulate) {
  const long long per_split = (long long)taps * cout * cin;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < per_split;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % cin);
    const int co = (int)((idx / cin) % cout);
    const int tap = (int)(idx / ((long long)cin * cout));
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[(long long)s * per_split + idx];
    float* o = dw + co * s_co + (long long)(ci_off + ci) * s_ci + tap * s_tap;
    *o = accumulate ? (*o + acc * scale) : acc * scale;
  }
}

// nn.Conv2d 3x3 layout ([Cout, Cin, 3, 3]: s_ci = 9, s_tap = 1): the generic kernel above writes with a 36-byte stride
// between neighbouring threads (one sector per 4-byte store). Here a block owns (co, 32 input channels): warp `tap`
// sums the splits of its tap with coalesced 128-byte reads (same split order as above: bitwise the same sums), the
// 9 x 32 results turn through shared memory and leave as 288 consecutive floats.
__global__ void __launch_bounds__(288)
wgrad_reduce_conv9_kernel(const float* __restrict__ partial, int splits, int cout, int cin, float scale,
                          float* __restrict__ dw, long long s_co, int ci_off, int accumulate) {
  __shared__ float tile[9][33];
  const int co = blockIdx.y, ci0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, tap = threadIdx.x >> 5;
  const long long per_split = 9LL * cout * cin;
  float acc = 0.f;
  if (ci0 + lane < cin) {
    const float* p = partial + ((long long)tap * cout + co) * cin + ci0 + lane;
    for (int s = 0; s < splits; ++s) acc += p[(long long)s * per_split];
  }
  tile[tap][lane] = acc * scale;
  __syncthreads();
  const int cl = threadIdx.x / 9, tp = threadIdx.x - cl * 9;   // output order: ci_local * 9 + tap
  if (ci0 + cl < cin) {
    float* o = dw + co * s_co + (long long)(ci_off + ci0 + cl) * 9 + tp;
    const float v = tile[tp][cl];
    *o = accumulate ? (*o + v) : v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Direct wgrad: MN-major operands straight from the NHWC tensors (no pixel-major copies)
// ---------------------------------------------------------------------------------------------------------
// The contraction index is the pixel, and in NHWC the channels of one pixel are contiguous, so a TMA box of
// [64 channels x 16 x 8 pixels] lands in shared memory as 128 rows (pixels = K) of 128 bytes (64 channels = M or N):
// exactly the MN-major SWIZZLE_128B canonical layout of tcgen05 (8 K-rows of 128 bytes per 1024-byte group, the next
// 64 channels LBO bytes further). One CTA owns (128 output channels) x (n_tile <= 128 input channels) x (the three kx
// taps of one ky) and loops over its share of 16x8-pixel tiles: per tile it loads the gradient tile and ONE
// (16+2)x(8+2)-pixel halo of the activation (TMA zero-fills outside the image = the convolution padding); the operand
// of tap (ky, kx) for tile row r is the 16 consecutive halo rows starting (r + ky) * 18 + kx rows into that copy - a
// descriptor start-address shift, valid because the swizzle is a function of absolute shared-memory address bits
// (same trick as the forward halo kernels). Three fp32 accumulators (one per kx) live in TMEM for the whole loop.
// Compared with the pixel-major path: no re-layout passes, and each loaded byte feeds 3 taps instead of 1.
constexpr int kDTW = 16, kDTH = 8;                       // pixel tile (16x4 with a 4-deep ring measured no faster at 64 px and 45% slower at 160 px)
constexpr int kDHW = kDTW + 2, kDHH = kDTH + 2;          // halo
constexpr int kDGBlock = kDTW * kDTH * 128;              // 16 KB: 128 pixels x 64 channels
constexpr int kDABlock = ((kDHW * kDHH * 128 + 1023) / 1024) * 1024;   // 23 KB (180 rows, padded to 1024)
constexpr int kDMaxStages = 4;

struct WgradDirectParams {
  int H, W, tiles_x, tiles_y, tiles_total, tiles_per_split;
  int taps, cout, cin, n_tile, n_blocks, a_cblocks;   // a_cblocks: 64-channel blocks of the activation per CTA (1 or 2)
  int g_coff, a_coff;
  int num_stages, tmem_cols;
  uint32_t stage_bytes;
  float* partial;   // [splits][taps][cout][cin]
};

// CLUSTER = true: the three ky CTAs of one (channel block, split) form a thread-block cluster (1 x 3 x 1). They walk the
// same tile sequence and need the SAME gradient tile and activation halo, so every box is fetched from L2 once and
// multicast into the three CTAs' shared memory (rank 0 issues the two gradient boxes, ranks 1 and 2 one activation box
// each); a stage is recycled when the MMAs of all three CTAs have consumed it (multicast tcgen05.commit on the three
// empty barriers). This cuts the L2 -> shared-memory traffic of the kernel by 3x - but measured slower than independent
// CTAs (see csd_wgrad_direct_bf16), so it is an opt-in experiment, kept correct by the same parity tests.
template <bool CLUSTER>
__global__ void __launch_bounds__(kWThreads, 1)
wgrad_direct_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapA,
                    const WgradDirectParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + p.num_stages * p.stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kDMaxStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kDMaxStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kDMaxStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mb = blockIdx.x / p.n_blocks, nb = blockIdx.x % p.n_blocks;
  const int m0 = mb * kWM, n0 = nb * p.n_tile;
  const int ky = p.taps == 9 ? (int)blockIdx.y : 1;
  const int ntap = p.taps == 9 ? 3 : 1;
  const int split = blockIdx.z;
  const int t_lo = split * p.tiles_per_split;
  const int t_hi = min(p.tiles_total, t_lo + p.tiles_per_split);
  const uint32_t a_off = 2u * kDGBlock;                  // activation blocks follow the two gradient blocks

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapG);
    ptx::prefetch_tensormap(&mapA);
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(full_bar(s), CLUSTER ? 1 : 2);
      ptx::mbar_init(empty_bar(s), CLUSTER ? 3 : 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (CLUSTER) ptx::cluster_sync();       // peers' barriers are initialised before anyone multicasts into them
  ptx::tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (CLUSTER && warp == 0) {
    if (lane == 0) {
      // ===== cluster producer: this CTA's share of every stage, multicast to the three ky CTAs =====
      const uint32_t rank = ptx::cluster_ctarank();
      const uint32_t total_bytes = 2u * kDGBlock + (uint32_t)p.a_cblocks * (kDHW * kDHH * 128);
      uint32_t stage = 0, par = 1;
      for (int t = t_lo; t < t_hi; ++t) {
        const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
        const int x0 = tx * kDTW, y0 = ty * kDTH;
        ptx::mbar_wait(empty_bar(stage), par);      // all three CTAs have consumed the previous use of this stage
        const uint32_t base = smem_base + stage * p.stage_bytes;
        ptx::mbar_arrive_expect_tx(full_bar(stage), total_bytes);
        if (rank == 0) {
          ptx::tma_load_4d_multicast(base, &mapG, full_bar(stage), p.g_coff + m0, x0, y0, b, 7);
          ptx::tma_load_4d_multicast(base + kDGBlock, &mapG, full_bar(stage), p.g_coff + m0 + 64, x0, y0, b, 7);
        } else if ((int)rank <= p.a_cblocks) {
          const int cb = (int)rank - 1;
          ptx::tma_load_4d_multicast(base + a_off + cb * kDABlock, &mapA, full_bar(stage), p.a_coff + n0 + cb * 64, x0 - 1,
                                     y0 - 1, b, 7);
        }
        if (++stage == (uint32_t)p.num_stages) { stage = 0; par ^= 1u; }
      }
    }
  } else if (!CLUSTER && (warp == 0 || warp == 6)) {
    if (lane == 0) {
      // ===== producers: warp 0 loads the gradient tile (M side), warp 6 the activation halo (N side) =====
      const bool is_g = warp == 0;
      uint32_t stage = 0, par = 1;
      for (int t = t_lo; t < t_hi; ++t) {
        const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
        const int x0 = tx * kDTW, y0 = ty * kDTH;
        ptx::mbar_wait(empty_bar(stage), par);
        const uint32_t base = smem_base + stage * p.stage_bytes;
        if (is_g) {
          ptx::mbar_arrive_expect_tx(full_bar(stage), 2u * kDGBlock);
          ptx::tma_load_4d(base, &mapG, full_bar(stage), p.g_coff + m0, x0, y0, b);
          ptx::tma_load_4d(base + kDGBlock, &mapG, full_bar(stage), p.g_coff + m0 + 64, x0, y0, b);
        } else {
          ptx::mbar_arrive_expect_tx(full_bar(stage), (uint32_t)p.a_cblocks * (kDHW * kDHH * 128));
          for (int cb = 0; cb < p.a_cblocks; ++cb)
            ptx::tma_load_4d(base + a_off + cb * kDABlock, &mapA, full_bar(stage), p.a_coff + n0 + cb * 64, x0 - 1, y0 - 1, b);
        }
        if (++stage == (uint32_t)p.num_stages) { stage = 0; par ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: both operands MN-major (a_major = b_major = 1) =====
      const uint32_t idesc = ptx::make_idesc_bf16_m128((uint32_t)p.n_tile) | (1u << 15) | (1u << 16);
      const uint32_t hi = ptx::smem_desc_hi(1024, kLayoutSw128);
      uint32_t stage = 0, par = 0, accumulate = 0;
      for (int t = t_lo; t < t_hi; ++t) {
        ptx::mbar_wait(full_bar(stage), par);
        ptx::tcgen05_fence_after();
        const uint32_t base = smem_base + stage * p.stage_bytes;
        const uint32_t g_lo = ptx::smem_desc_lo(base, kDGBlock);             // LBO = next 64 output channels
        const uint32_t a_lo = ptx::smem_desc_lo(base + a_off, kDABlock);     // LBO = next 64 input channels
#pragma unroll 1
        for (int r = 0; r < kDTH; ++r) {
          const uint32_t g_r = g_lo + ((r * kDTW * 128) >> 4);
          for (int kx = 0; kx < ntap; ++kx) {
            const int dx = p.taps == 9 ? kx : 1;
            const uint32_t a_r = a_lo + ((((r + ky) * kDHW + dx) * 128) >> 4);
            ptx::mma_bf16_ss(tmem_base + kx * p.n_tile, ptx::smem_desc_join(hi, g_r), ptx::smem_desc_join(hi, a_r), idesc,
                             accumulate);
          }
          accumulate = 1u;
        }
        if (CLUSTER) ptx::mma_commit_multicast(empty_bar(stage), 7);
        else ptx::mma_commit(empty_bar(stage));
        if (++stage == (uint32_t)p.num_stages) { stage = 0; par ^= 1u; }
      }
      ptx::mma_commit(tmem_full_bar);
    }
  } else if (warp >= 2 && warp <= 5) {
    // ===== epilogue (warps 2..5): TMEM lane = output channel, one accumulator per kx =====
    const int q = warp & 3;
    const int co = m0 + q * 32 + lane;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tcgen05_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ncols = min(p.n_tile, p.cin - n0);
    for (int kx = 0; kx < ntap; ++kx) {
      const int tap = p.taps == 9 ? ky * 3 + kx : 0;
      float* orow = p.partial + (((long long)split * p.taps + tap) * p.cout + co) * p.cin;
      for (int col = 0; col < ncols; col += 16) {
        uint32_t r[16];
        __syncwarp();
        ptx::tmem_ld_x16(t_row + kx * p.n_tile + col, r);
        ptx::tmem_ld_wait();
        if (co < p.cout) store_partial_chunk(orow + n0 + col, r, min(16, ncols - col), !(t_hi > t_lo));
      }
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (CLUSTER) ptx::cluster_sync();       // no CTA leaves while a peer may still multicast into it or signal its barriers
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// input channels per CTA of the direct kernel (developer knob CSD_WGRAD_NMAX = 64 trades operand re-reads for a deeper ring)
static int direct_n_max() {
  static const int v = [] {
    const char* e = getenv("CSD_WGRAD_NMAX");
    const int n = e ? atoi(e) : 128;
    return (n == 64 || n == 128) ? n : 128;
  }();
  return v;
}

static int next_pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

}  // namespace csd

extern "C" {

int csd_pixmajor_geometry(int batch, int grid_h, int grid_w, csd_pixmajor_geom* g) {
  using namespace csd;
  CSD_REQUIRE(g != nullptr && batch >= 1 && grid_h >= 1 && grid_w >= 1, "pixmajor_geometry: bad arguments");
  g->wp = (grid_w + 2 + 7) / 8 * 8;
  g->q = (int64_t)(grid_h + 2) * g->wp;
  // images per split: K per CTA of at least ~4096 pixels, but keep >= ~16 splits for the large images
  int64_t ips = (4096 + g->q - 1) / g->q;
  if (ips < 1) ips = 1;
  if (ips > batch) ips = batch;
  g->ips = (int32_t)ips;
  g->splits = (batch + g->ips - 1) / g->ips;
  const int64_t ks = g->ips * g->q;
  g->kp = (ks + 63) / 64 * 64;
  g->margin = (g->wp + 8 + 63) / 64 * 64;
  g->row_pitch = g->margin + g->kp + g->margin;
  return CSD_OK;
}

int csd_nhwc_to_pixmajor_bf16(const void* src, int pitch, int c_off, int c_cnt, int batch, int h, int w, int stride,
                              int offset, const csd_pixmajor_geom* g, int ncopies, void* out, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(src && out && g, "nhwc_to_pixmajor: null pointer");
  CSD_REQUIRE(ncopies == 1 || ncopies == 3, "nhwc_to_pixmajor: ncopies %d (1 or 3)", ncopies);
  CSD_REQUIRE(c_cnt >= 1 && c_off >= 0 && c_off + c_cnt <= pitch, "nhwc_to_pixmajor: bad channel range");
  CSD_REQUIRE(stride >= 1 && offset >= 0 && batch >= 1 && batch <= 65535 && h >= 1 && h <= 65535 && w >= 1,
              "nhwc_to_pixmajor: bad shape");
  // the last source pixel must land inside the padded grid the geometry was built for
  CSD_REQUIRE((w - 1) * stride + offset + 2 < g->wp && (int64_t)((h - 1) * stride + offset + 2) * g->wp <= g->q,
              "nhwc_to_pixmajor: source %dx%d (stride %d, offset %d) does not fit the grid (wp=%d)", h, w, stride, offset,
              g->wp);
  PixMajorParams p;
  p.src = static_cast<const __nv_bfloat16*>(src);
  p.pitch = pitch; p.c_off = c_off; p.c_cnt = c_cnt;
  p.batch = batch; p.h = h; p.w = w; p.stride = stride; p.offset = offset;
  p.wp = g->wp; p.q = g->q; p.ips = g->ips; p.margin = g->margin; p.row_pitch = g->row_pitch;
  p.ncopies = ncopies;
  p.copy_stride = (long long)g->splits * c_cnt * g->row_pitch;
  p.out = static_cast<__nv_bfloat16*>(out);
  dim3 grid((unsigned)ceil_div(c_cnt, 64), (unsigned)h, (unsigned)batch);
  pixmajor_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CSD_LAUNCH_CHECK("pixmajor_kernel");
  return CSD_OK;
}

int csd_wgrad_gemm_bf16(const void* g_pm, int cout, const void* a_pm, int cin, int taps, const csd_pixmajor_geom* g,
                        float* partial, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(g_pm && a_pm && g && partial, "wgrad_gemm: null pointer");
  CSD_REQUIRE(taps == 1 || taps == 9, "wgrad_gemm: taps %d (1 or 9)", taps);
  CSD_REQUIRE(cout >= 1 && cin >= 1, "wgrad_gemm: bad channel counts");
  WgradParams p;
  p.taps = taps; p.wp = g->wp; p.margin = g->margin; p.k_chunks = (int)(g->kp / kWK);
  p.cout = cout; p.cin = cin;
  const int n16 = ceil_div(cin, 16) * 16;
  p.n_tiles = ceil_div(n16, 256);
  p.n_tile = ceil_div(ceil_div(n16, p.n_tiles), 16) * 16;
  p.b_bytes = (uint32_t)p.n_tile * kWK * 2;
  p.stage_bytes = kWABytes + p.b_bytes;
  p.num_stages = std::min<int>(kWMaxStages, (int)((200 * 1024) / p.stage_bytes));
  p.tmem_cols = next_pow2_cols(p.n_tile);
  p.partial = partial;
  const int m_tiles = ceil_div(cout, kWM);

  CUtensorMap mapG, mapA[3];
  {
    uint64_t dims[3] = {(uint64_t)g->row_pitch, (uint64_t)cout, (uint64_t)g->splits};
    uint64_t strides[2] = {(uint64_t)g->row_pitch * 2, (uint64_t)g->row_pitch * 2 * cout};
    uint32_t box[3] = {(uint32_t)kWK, (uint32_t)kWM, 1};
    int st = encode_tensor_map(&mapG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, g_pm, dims, strides, box, TMA_SW_128);
    if (st != CSD_OK) return st;
  }
  const long long copy_stride = (long long)g->splits * cin * g->row_pitch;
  for (int kx = 0; kx < 3; ++kx) {
    // a 1-tap contraction has the centre copy only (stored first)
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(a_pm) + (taps == 9 ? kx * copy_stride : 0);
    uint64_t dims[3] = {(uint64_t)g->row_pitch, (uint64_t)cin, (uint64_t)g->splits};
    uint64_t strides[2] = {(uint64_t)g->row_pitch * 2, (uint64_t)g->row_pitch * 2 * cin};
    uint32_t box[3] = {(uint32_t)kWK, (uint32_t)p.n_tile, 1};
    int st = encode_tensor_map(&mapA[kx], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, TMA_SW_128);
    if (st != CSD_OK) return st;
  }
  const size_t smem = (size_t)p.num_stages * p.stage_bytes + 8 * (2 * kWMaxStages + 2) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    CSD_CUDA(cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)(m_tiles * p.n_tiles), (unsigned)taps, (unsigned)g->splits);
  wgrad_gemm_kernel<<<grid, kWThreads, smem, static_cast<cudaStream_t>(stream)>>>(mapG, mapA[0], mapA[1], mapA[2], p);
  CSD_LAUNCH_CHECK("wgrad_gemm_kernel");
  return CSD_OK;
}

int csd_wgrad_direct_splits(int batch, int h, int w, int cout, int cin, int taps, int* splits) {
  using namespace csd;
  CSD_REQUIRE(splits != nullptr && batch >= 1 && h >= 1 && w >= 1 && cout >= 1 && cin >= 1, "wgrad_direct_splits: bad arguments");
  const int tiles = batch * ceil_div(h, kDTH) * ceil_div(w, kDTW);
  const int n_blocks = ceil_div(cin, direct_n_max());
  const int ctas = ceil_div(cout, kWM) * n_blocks * (taps == 9 ? 3 : 1);
  int s = ceil_div(2 * num_sms(), ctas);
  s = std::max(1, std::min(s, std::max(1, tiles / 4)));
  const int tps = ceil_div(tiles, s);
  *splits = ceil_div(tiles, tps);
  return CSD_OK;
}

int csd_wgrad_direct_bf16(const void* g, int g_pitch, int g_c_off, int cout, const void* a, int a_pitch, int a_c_off,
                          int cin, int taps, int batch, int h, int w, float* partial, int splits, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(g && a && partial, "wgrad_direct: null pointer");
  CSD_REQUIRE(taps == 1 || taps == 9, "wgrad_direct: taps %d (1 or 9)", taps);
  CSD_REQUIRE(g_pitch % 8 == 0 && a_pitch % 8 == 0 && g_c_off % 8 == 0 && a_c_off % 8 == 0,
              "wgrad_direct: channel pitches / offsets must be multiples of 8");
  CSD_REQUIRE(g_pitch >= 64 && a_pitch >= 64, "wgrad_direct: tensors with fewer than 64 channels take the pixel-major path");
  WgradDirectParams p;
  p.H = h; p.W = w;
  p.tiles_x = ceil_div(w, kDTW); p.tiles_y = ceil_div(h, kDTH);
  p.tiles_total = batch * p.tiles_x * p.tiles_y;
  int want = 0;
  int st = csd_wgrad_direct_splits(batch, h, w, cout, cin, taps, &want);
  if (st != CSD_OK) return st;
  CSD_REQUIRE(splits == want, "wgrad_direct: splits %d != csd_wgrad_direct_splits() = %d", splits, want);
  p.tiles_per_split = ceil_div(p.tiles_total, splits);
  p.taps = taps; p.cout = cout; p.cin = cin;
  p.n_blocks = ceil_div(cin, direct_n_max());
  p.n_tile = ceil_div(ceil_div(cin, p.n_blocks), 16) * 16;
  p.a_cblocks = ceil_div(p.n_tile, 64);
  p.g_coff = g_c_off; p.a_coff = a_c_off;
  p.stage_bytes = 2u * kDGBlock + (uint32_t)p.a_cblocks * kDABlock;
  p.num_stages = std::min<int>(kDMaxStages, (int)((200 * 1024) / p.stage_bytes));
  p.tmem_cols = next_pow2_cols((taps == 9 ? 3 : 1) * p.n_tile);
  p.partial = partial;
  CUtensorMap mapG, mapA;
  {
    uint64_t dims[4] = {(uint64_t)g_pitch, (uint64_t)w, (uint64_t)h, (uint64_t)batch};
    uint64_t strides[3] = {(uint64_t)g_pitch * 2, (uint64_t)g_pitch * 2 * w, (uint64_t)g_pitch * 2 * w * h};
    uint32_t box[4] = {64, (uint32_t)kDTW, (uint32_t)kDTH, 1};
    st = encode_tensor_map(&mapG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, g, dims, strides, box, TMA_SW_128);
    if (st != CSD_OK) return st;
  }
  {
    uint64_t dims[4] = {(uint64_t)a_pitch, (uint64_t)w, (uint64_t)h, (uint64_t)batch};
    uint64_t strides[3] = {(uint64_t)a_pitch * 2, (uint64_t)a_pitch * 2 * w, (uint64_t)a_pitch * 2 * w * h};
    uint32_t box[4] = {64, (uint32_t)kDHW, (uint32_t)kDHH, 1};
    st = encode_tensor_map(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a, dims, strides, box, TMA_SW_128);
    if (st != CSD_OK) return st;
  }
  const size_t smem = (size_t)p.num_stages * p.stage_bytes + 8 * (2 * kDMaxStages + 2) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    CSD_CUDA(cudaFuncSetAttribute(wgrad_direct_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CSD_CUDA(cudaFuncSetAttribute(wgrad_direct_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)(ceil_div(cout, kWM) * p.n_blocks), (unsigned)(taps == 9 ? 3 : 1), (unsigned)splits);
  // Opt-in (CSD_WGRAD_CLUSTER=1): measured on B200 the multicast variant is SLOWER than three independent CTAs (64 px,
  // 128 -> 128 channels, batch 50: 205 us vs 146 us; 160 px, 96 -> 96, batch 64: 749 us vs 515 us, tools/wgrad_bench.py):
  // the kernel is not L2-bandwidth bound, and coupling three CTAs per 2-deep stage ring costs more than it saves.
  static const bool use_cluster = getenv("CSD_WGRAD_CLUSTER") != nullptr;
  if (taps == 9 && use_cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kWThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 3;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CSD_CUDA(cudaLaunchKernelEx(&cfg, wgrad_direct_kernel<true>, mapG, mapA, p));
    count_launch();
    return CSD_OK;
  }
  wgrad_direct_kernel<false><<<grid, kWThreads, smem, static_cast<cudaStream_t>(stream)>>>(mapG, mapA, p);
  CSD_LAUNCH_CHECK("wgrad_direct_kernel");
  return CSD_OK;
}

int csd_wgrad_reduce_f32(const float* partial, int splits, int taps, int cout, int cin, float scale, float* dw,
                         int64_t stride_co, int64_t stride_ci, int64_t stride_tap, int ci_off, int accumulate,
                         csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(partial && dw && splits >= 1 && taps >= 1 && cout >= 1 && cin >= 1, "wgrad_reduce: bad arguments");
  const long long per_split = (long long)taps * cout * cin;
  if (taps == 9 && stride_tap == 1 && stride_ci == 9 && cout <= 65535) {
    wgrad_reduce_conv9_kernel<<<dim3((unsigned)ceil_div(cin, 32), (unsigned)cout), 288, 0,
                                static_cast<cudaStream_t>(stream)>>>(partial, splits, cout, cin, scale, dw, stride_co, ci_off,
                                                                     accumulate);
    CSD_LAUNCH_CHECK("wgrad_reduce_conv9_kernel");
    return CSD_OK;
  }
  const int blocks = (int)std::min<long long>(ceil_div_ll(per_split, 256), (long long)num_sms() * 8);
  wgrad_reduce_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(partial, splits, taps, cout, cin, scale, dw,
                                                                           stride_co, stride_ci, stride_tap, ci_off,
                                                                           accumulate);
  CSD_LAUNCH_CHECK("wgrad_reduce_kernel");
  return CSD_OK;
}

}  // extern "C"
