// tcgen05 implicit-GEMM convolution / GEMM for NHWC bf16 activations (sm_100a).
//
// One CTA computes a 128-pixel x n_tile output tile:
//   warp 0   : TMA producer   - per (segment, tap, 32-channel chunk) one 4-D box load of the shifted
//                               activation tile (out-of-image pixels are zero-filled by TMA = conv
//                               padding) and one 3-D box load of the matching weight slab
//   warp 1   : MMA issuer     - tcgen05.mma (M=128, N=n_sub<=256, K=16) from shared-memory
//                               descriptors, fp32 accumulators in TMEM; tcgen05.commit releases stages
//   warps 2-5: epilogue       - tcgen05.ld the accumulator rows, add bias / time-embedding
//                               projection / residual, scale, store bf16 or fp32 rows
// Shared memory holds a ring of `num_stages` {A: 128 rows x 64 B, B: n_tile rows x 64 B} stages in
// the 64-byte-swizzled K-major layout shared by TMA and the UMMA descriptors.
//
// Reference ops replaced: nn.Conv2d 3x3/1x1 (models/layers.py:100-132), NIN (models/layers.py:555-564),
// attention einsums (models/layerspp.py:82-86), `h += Dense_0(act(temb))[:, :, None, None]`, the
// Conv_2 skip and `(x + h) / sqrt(2)` of ResnetBlockBigGANpp.forward (models/layerspp.py:260-274).
#include "common.cuh"
#include "ptx.cuh"
#include "tensormap.cuh"
#include "../../include/csd_b200.h"

#include <algorithm>
#include <utility>
#include <vector>
#include <cstdlib>

namespace csd {

#ifdef CSD_ENABLE_PHASE_TIMESTAMPS
#define CSD_TS(i)                                                                          \
  do {                                                                                   \
    if (p.debug_ts != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0) \
      p.debug_ts[i] = clock64();                                                         \
  } while (0)
// persistent kernel: timestamps of the probe CTA's 6th tile, one slot per role event
#define CSD_TSP(i)                                                                          \
  do {                                                                                    \
    if (p.debug_ts != nullptr && blockIdx.x == gridDim.x / 2 && tile == (int)blockIdx.x + 5 * (int)gridDim.x) \
      p.debug_ts[i] = clock64();                                                          \
  } while (0)
#else
#define CSD_TS(i) do { } while (0)
#define CSD_TSP(i) do { } while (0)
#endif

constexpr int kChunkK = 32;                       // channels per pipeline stage
constexpr int kRowBytes = kChunkK * 2;            // 64 B of bf16 per row -> SWIZZLE_64B
constexpr int kTileM = 128;                       // UMMA M
constexpr int kAStageBytes = kTileM * kRowBytes;  // 8192
constexpr int kConvThreads = 192;
// conv_gemm_kernel: a single thread can start one TMA load every ~360 cycles however small the box is (measured,
// tools/tma_rate_probe.cu: 8 KB and 16 KB boxes, tiled or bulk, all cost the same per issue; two issuing threads
// reach 1.7x), while one pipeline stage is only ~200 tensor cycles of work. The loads of consecutive stages are
// therefore issued by kTapProducers threads in different warps, round-robin.
constexpr int kTapProducers = 4;
constexpr int kTapThreads = kConvThreads + 32 * (kTapProducers - 1);   // warps 6.. are the extra producers
constexpr int kMaxStages = 12;
constexpr uint32_t kLayoutSw64 = 4;

constexpr int kMaxSched = 256;   // chunk schedule entries of the transposed kernel (more chunks: segment order)

struct ConvGemmKernelParams {
  int B, H, W, TW, TH, TB;
  int tiles_w, tiles_h;
  int stride, pad;
  int nseg;
  int seg_taps[CSD_MAX_SEGMENTS];
  int seg_chunks[CSD_MAX_SEGMENTS];
  int seg_coff[CSD_MAX_SEGMENTS];
  int n_store, n_tile, n_sub, nsplit;
  int wt_k_off;
  int a_batch_step;
  int num_stages, tmem_cols;
  // transposed kernel: the order in which the (segment, chunk) pairs of a tile are walked - the same for the halo
  // producer, the weight producers, the transform warps and the MMA issuer. Short 1-tap chunks (two MMAs per pixel
  // halo: skip 1x1 convolutions, identity residuals) are spread between the 9-tap chunks (18 MMAs), so their TMA
  // latency hides behind tensor work instead of draining the 3-4 deep halo ring at the end of every tile.
  int staging_bytes;                     // epilogue staging tile of this launch (1 KB aligned)
  int b_box_rows;                        // weight rows per slab that TMA actually writes (<= 128; see conv_gemm_prepare)
  int n_sched;                           // chunks per tile
  int sched_tab;                         // 1: sched_seg / sched_chunk hold the order; 0: segment order (too many chunks)
  unsigned char sched_seg[kMaxSched], sched_chunk[kMaxSched];
  int seg_tab_base[CSD_MAX_SEGMENTS];    // first (scale, shift) slot of the segment in the coefficient table
  int ksplit;     // per-tap kernel: split-K over gridDim.z (each z computes a contiguous range of the (segment, tap, chunk)
                  // iterations and writes raw fp32 partial sums; csd_splitk_reduce_bf16 finishes the epilogue)
  int producers;  // per-tap kernel: issuing threads; num_stages is a multiple of it, so a ring slot always belongs to
                  // the same producer (the two-phase mbarrier parity scheme needs every use of a slot seen in order)
  uint32_t stage_bytes, a_box_bytes, b_box_bytes;
  float* stat_partials;  // [tiles, n_store, 2] per-tile per-channel (sum, sumsq) of the stored output, or null
  long long* debug_ts;  // perf experiment only: phase timestamps of one mid-grid CTA
  int debug_nodata;  // perf experiment only: skip the steady-state loads (results are garbage)
  // halo mode
  int mt, a_stages, b_stages;
  int seg_kbase[CSD_MAX_SEGMENTS];
  uint32_t a_stage_bytes, b_stage_bytes;
  // transposed mode: fused GroupNorm(+SiLU) prologue (per-segment (scale, shift) tables [batch, c_cnt, 2])
  int num_tiles, n_blocks;   // persistent transposed kernel: tiles = spatial tiles x 128-channel blocks
  int out_box_c;             // channel extent of its TMA store box = staging row pitch
  int t_rows, t_pix;         // macro tile = t_rows x 8 pixels (32 -> N = 256, or 20 -> N = 160 for 40-row images)
  const float* seg_norm[CSD_MAX_SEGMENTS];
  int seg_silu[CSD_MAX_SEGMENTS];
  int seg_ccnt[CSD_MAX_SEGMENTS];
  int has_norm;
  void* out;
  int out_pitch, out_f32;
  int out_round;   // fp32 output rounded to tf32 (tensors that only feed further tensor-core operands: q|k, V^T, P V)
  long long out_z_stride;
  const float* bias;
  int bias_per_row;
  const float* temb;
  int temb_pitch;
  const __nv_bfloat16* res;
  int res_pitch;
  long long res_z_stride;
  float scale;
};

// Epilogue of one accumulator row: thread = one output pixel (TMEM lane), 16 columns at a time.
// out = (acc + bias + temb[b] + residual) * scale, stored as bf16 or fp32 rows.
// Latency hiding: the per-column addend (bias + temb) comes from a shared-memory table built while the
// main loop runs (s_add, may be null), and the residual of chunk k+1 is fetched while chunk k's
// tcgen05.ld is in flight, so no global-load latency sits between TMEM read and the store.
template <bool kF32>
__device__ __forceinline__ void epilogue_rows(const ConvGemmKernelParams& p, uint32_t t_row, bool valid, long long pix,
                                              int b, int n0, int z, const float* s_add) {
  const float row_bias = (p.bias != nullptr && p.bias_per_row && valid) ? p.bias[pix] : 0.0f;
  const float* temb_row = (p.temb != nullptr && valid) ? p.temb + (long long)b * p.temb_pitch : nullptr;
  const __nv_bfloat16* res_row =
      (!kF32 && p.res != nullptr && valid) ? p.res + (long long)z * p.res_z_stride + pix * p.res_pitch : nullptr;
  // fp32-activation plan: the residual is an fp32 tensor addressed like the output
  const float* res_row_f = (kF32 && p.res != nullptr && valid)
                               ? reinterpret_cast<const float*>(p.res) + (long long)z * p.res_z_stride + pix * p.res_pitch
                               : nullptr;
  const int ncols = min(p.n_tile, p.n_store - n0);  // columns of this tile that are stored
  uint4 rn0 = make_uint4(0, 0, 0, 0), rn1 = rn0;
  if (res_row != nullptr && ncols >= 16) {
    const uint4* rp = reinterpret_cast<const uint4*>(res_row + n0);
    rn0 = __ldg(rp);
    rn1 = __ldg(rp + 1);
  }
  for (int col = 0; col < ncols; col += 16) {
    uint32_t r[16];
    __syncwarp();
    ptx::tmem_ld_x16(t_row + col, r);
    const int n = n0 + col;
    const int cnt = min(16, p.n_store - n);
    const uint4 rc0 = rn0, rc1 = rn1;
    if (res_row != nullptr && col + 32 <= ncols) {  // next chunk is complete: prefetch its residual
      const uint4* rp = reinterpret_cast<const uint4*>(res_row + n + 16);
      rn0 = __ldg(rp);
      rn1 = __ldg(rp + 1);
    }
    ptx::tmem_ld_wait();
    if (valid) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
      if (s_add != nullptr) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 a = *reinterpret_cast<const float4*>(s_add + col + i);
          v[i] += a.x; v[i + 1] += a.y; v[i + 2] += a.z; v[i + 3] += a.w;
        }
      } else {
        if (p.bias != nullptr) {
          if (p.bias_per_row) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += row_bias;
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + n + i);
          }
        }
        if (temb_row != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(temb_row + n + i);
        }
      }
      if (res_row != nullptr) {
        if (cnt == 16) {
          bf16x8 r0, r1;
          *reinterpret_cast<uint4*>(&r0) = rc0;
          *reinterpret_cast<uint4*>(&r1) = rc1;
          float f[16];
          unpack8(r0, f);
          unpack8(r1, f + 8);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += f[i];
        } else {
          // (static indices under a predicate: a dynamically indexed v[] would live in local memory and drag every
          //  access of the fast path through the stack as well)
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < cnt) v[i] += __bfloat162float(res_row[n + i]);
        }
      }
      if (kF32 && res_row_f != nullptr) {
        if (cnt == 16) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(res_row_f + n + i));
            v[i] += a.x; v[i + 1] += a.y; v[i + 2] += a.z; v[i + 3] += a.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < cnt) v[i] += __ldg(res_row_f + n + i);
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] *= p.scale;
      if (kF32 && p.out_round) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = round_tf32(v[i]);
      }

      if (kF32 || p.out_f32) {
        float* op = reinterpret_cast<float*>(p.out) + (long long)z * p.out_z_stride + pix * p.out_pitch + n;
        if (cnt == 16) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(op)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < cnt) op[i] = v[i];
        }
      } else {
        __nv_bfloat16* op =
            reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)z * p.out_z_stride + pix * p.out_pitch + n;
        if (cnt == 16) {
          bf16x8 o0 = pack8(v), o1 = pack8(v + 8);
          reinterpret_cast<uint4*>(op)[0] = *reinterpret_cast<uint4*>(&o0);
          reinterpret_cast<uint4*>(op)[1] = *reinterpret_cast<uint4*>(&o1);
        } else if (cnt == 8) {
          bf16x8 o0 = pack8(v);
          reinterpret_cast<uint4*>(op)[0] = *reinterpret_cast<uint4*>(&o0);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < cnt) op[i] = __float2bfloat16_rn(v[i]);
        }
      }
    }
  }
}

// Built by the 128 epilogue threads while the main loop runs: s_add[bl][i] = bias[n0+i] + temb[b0+bl, n0+i] for the
// tile's tb images (tb > 1 at the <= 8 px levels, where one 128-pixel tile spans several images). Only when the
// bias is per column and the table fits its kAddendFloats slots.
constexpr int kAddendFloats = 2048;
__device__ __forceinline__ const float* build_addend_table(const ConvGemmKernelParams& p, uint32_t s_add_addr, int b0,
                                                           int n0, int et /*0..127*/) {
  if (p.bias_per_row || (p.bias == nullptr && p.temb == nullptr) || p.TB * p.n_tile > kAddendFloats) return nullptr;
  float* s_add = reinterpret_cast<float*>(__cvta_shared_to_generic(s_add_addr));
  for (int idx = et; idx < p.TB * p.n_tile; idx += 128) {
    const int bl = idx / p.n_tile, i = idx - bl * p.n_tile;
    float a = p.bias != nullptr ? __ldg(p.bias + n0 + i) : 0.f;
    if (p.temb != nullptr && b0 + bl < p.B) a += __ldg(p.temb + (long long)(b0 + bl) * p.temb_pitch + n0 + i);
    s_add[idx] = a;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
  return s_add;
}

// kChunk = channels per pipeline stage: 64 (128-byte rows, SWIZZLE_128B, 4 MMAs per stage) for every launch with a
// segment wider than 32 channels, 32 (64-byte rows, SWIZZLE_64B, 2 MMAs) for the narrow ones (6-channel input conv).
// The single issuing thread spends ~450 cycles of scalar bookkeeping per stage (measured with the loads switched off,
// tools/tap_nodata_probe.py: the K loop takes the same time with and without data), so the stage has to carry
// more tensor work than that; the weight K layout stays taps x ceil32(C): a 64-channel box that runs past the end
// of a tap reads the next tap's columns against activation channels the TMA unit zero-fills.
template <int kChunk, bool kTf32>
__global__ void __launch_bounds__(kTapThreads)
conv_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                 const __grid_constant__ CUtensorMap mapB, const ConvGemmKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kRowB = kChunk * (kTf32 ? 4 : 2);             // bytes per operand row of a stage: 128 or 64
  constexpr uint32_t kARows = kTileM * kRowB;                      // bytes of the A part of a stage
  constexpr uint32_t kLayout = kRowB == 128 ? 2u : kLayoutSw64;    // UMMA layout type: SWIZZLE_128B / SWIZZLE_64B
  constexpr uint32_t kSbo = 8u * kRowB;                            // 8 rows of the swizzle atom
  // 1024-byte aligned base: the swizzle pattern is a function of the shared-memory address bits.
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + p.num_stages * p.stage_bytes;
  // barrier layout: full[num_stages], empty[num_stages], tmem_full, then the TMEM address slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kMaxStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 1);
  const uint32_t s_add_addr = bar_base + 8u * (2 * kMaxStages + 2);  // 16-byte aligned, n_tile floats

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int t = blockIdx.x;
  const int tw = t % p.tiles_w;
  const int th = (t / p.tiles_w) % p.tiles_h;
  const int tb = t / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.TW, h0 = th * p.TH, b0 = tb * p.TB;
  const int n0 = blockIdx.y * p.n_tile;
  const int z = blockIdx.z;

  int all_iters = 0;
  for (int s = 0; s < p.nseg; ++s) all_iters += p.seg_taps[s] * p.seg_chunks[s];
  // split-K: this CTA owns iterations [it_lo, it_hi) of the flattened (segment, tap, chunk) list
  const int it_lo = p.ksplit > 1 ? (int)(((long long)all_iters * z) / p.ksplit) : 0;
  const int it_hi = p.ksplit > 1 ? (int)(((long long)all_iters * (z + 1)) / p.ksplit) : all_iters;
  const int total_iters = it_hi - it_lo;
  const int zb = p.ksplit > 1 ? 0 : z;              // batched-GEMM index of the operands (split-K shares them)
  if (threadIdx.x == 0) CSD_TS(0);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapA0);
    ptx::prefetch_tensormap(&mapB);
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
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
  if (threadIdx.x == 0) CSD_TS(1);

  // Single-thread issue loops: ring position / parity are running counters (no division) and the UMMA
  // descriptors advance by adding to their low word, so the instruction stream between two tcgen05.mma
  // stays shorter than the MMA itself.
  if (warp == 0 || warp >= kConvThreads / 32) {
    // ===== TMA producers: producer j issues the loads of the stages whose running index is j mod kTapProducers =====
    if (lane == 0) {
      const int me = (warp == 0) ? 0 : warp - kConvThreads / 32 + 1;
      uint32_t stage = 0, par = 1;
      int turn = 0;
      int kcol = p.wt_k_off;  // running K column into Wt
      int issued = 0;         // perf experiment (debug_nodata): loads beyond the first ring fill can be switched off
      int git = 0;            // index in the flattened (segment, tap, chunk) list
      for (int s = 0; s < p.nseg; ++s) {
        const CUtensorMap* mapA = (s == 0) ? &mapA0 : (s == 1) ? &mapA1 : (s == 2) ? &mapA2 : &mapA3;
        const int taps = p.seg_taps[s];
        const int kpad = ((p.seg_ccnt[s] + kChunkK - 1) / kChunkK) * kChunkK;   // K columns of one tap in Wt
        for (int tap = 0; tap < taps; ++tap, kcol += kpad) {
          const int dy = (taps == 9) ? (tap / 3 - p.pad) : 0;
          const int dx = (taps == 9) ? (tap % 3 - p.pad) : 0;
          const int cw = w0 * p.stride + dx, ch = h0 * p.stride + dy;
          for (int c = 0; c < p.seg_chunks[s]; ++c, ++git) {
            if (git < it_lo || git >= it_hi) continue;      // another split's iteration
            if (turn == me) {
              ptx::mbar_wait(empty_bar(stage), par);
#ifdef CSD_ENABLE_PHASE_TIMESTAMPS
              if (p.debug_ts != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0 && me == 0 &&
                  issued < 96)
                p.debug_ts[64 + issued] = clock64();
#endif
              const uint32_t a_dst = smem_base + stage * p.stage_bytes;
              const uint32_t b_dst = a_dst + kARows;
              const int kc = kcol + c * kChunk;
              const bool steady = p.debug_nodata != 0 && issued >= p.num_stages;
              const bool load_a = !(steady && (p.debug_nodata & 2)), load_b = !(steady && (p.debug_nodata & 1));
              ptx::mbar_arrive_expect_tx(full_bar(stage),
                                         (load_a ? p.a_box_bytes : 0u) + (load_b ? p.b_box_bytes * p.nsplit : 0u));
              if (load_a)
                ptx::tma_load_4d(a_dst, mapA, full_bar(stage), p.seg_coff[s] + c * kChunk, cw, ch, b0 + zb * p.a_batch_step);
              if (load_b) {
                ptx::tma_load_3d(b_dst, &mapB, full_bar(stage), kc, n0, zb);
                if (p.nsplit > 1) ptx::tma_load_3d(b_dst + p.b_box_bytes, &mapB, full_bar(stage), kc, n0 + p.n_sub, zb);
              }
            }
            ++issued;
            if (++turn == p.producers) turn = 0;
            if (++stage == (uint32_t)p.num_stages) { stage = 0; par ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the loop converged, one elected lane issues (see conv_halo_tp_kernel) =====
    {
      // everything the loop needs lives in registers: the issuing thread's scalar work per stage is what bounds
      // the small-level launches
      const uint32_t idesc = kTf32 ? ptx::make_idesc_tf32_m128((uint32_t)p.n_sub) : ptx::make_idesc_bf16_m128((uint32_t)p.n_sub);
      const uint32_t hi = ptx::smem_desc_hi(kSbo, kLayout);
      const bool split = p.nsplit > 1;
      const uint32_t b2_off = p.b_box_bytes >> 4;
      const uint32_t num_stages = (uint32_t)p.num_stages, stage_bytes = p.stage_bytes;
      const uint32_t tmem_d0 = tmem_base, tmem_d1 = tmem_base + (uint32_t)p.n_sub;
      uint32_t stage = 0, par = 0, accumulate = 0;
      uint32_t a_lo = ptx::smem_desc_lo(smem_base, 16);
      const uint32_t a_lo0 = a_lo, stage_step = stage_bytes >> 4;
      uint32_t fbar = full_bar(0);
      bool ready = ptx::mbar_test_wait(fbar, 0);
#ifdef CSD_ENABLE_PHASE_TIMESTAMPS
      const bool ts_on = lane == 0 && p.debug_ts != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0;
#define CSD_TSM(slot) do { if (ts_on && it < 12) p.debug_ts[16 + 4 * it + (slot)] = clock64(); } while (0)
#else
#define CSD_TSM(slot) do { } while (0)
#endif
      for (int it = 0; it < total_iters; ++it) {
        CSD_TSM(0);
        if (!ready) ptx::mbar_wait(fbar, par);
        ptx::tcgen05_fence_after();
        CSD_TSM(1);
        // look at the next stage's barrier now: the test's latency overlaps with the MMAs issued below
        const uint32_t ebar = fbar + 8u * kMaxStages;
        uint32_t nstage = stage + 1, npar = par, n_lo = a_lo + stage_step;
        fbar += 8u;
        if (nstage == num_stages) { nstage = 0; npar ^= 1u; n_lo = a_lo0; fbar = full_bar(0); }
        ready = (it + 1 < total_iters) && ptx::mbar_test_wait(fbar, npar);
        const uint32_t b_lo = a_lo + (kARows >> 4);
        if (ptx::elect_one()) {
#pragma unroll
        for (int k16 = 0; k16 < (int)kRowB / 32; ++k16) {      // one MMA per 32 bytes of K: 16 bf16 or 8 tf32 values
          const uint32_t acc = (k16 == 0) ? accumulate : 1u;
          if (kTf32) {
            ptx::mma_tf32_ss(tmem_d0, ptx::smem_desc_join(hi, a_lo + 2 * k16), ptx::smem_desc_join(hi, b_lo + 2 * k16),
                             idesc, acc);
            if (split)
              ptx::mma_tf32_ss(tmem_d1, ptx::smem_desc_join(hi, a_lo + 2 * k16),
                               ptx::smem_desc_join(hi, b_lo + b2_off + 2 * k16), idesc, acc);
          } else {
            ptx::mma_bf16_ss(tmem_d0, ptx::smem_desc_join(hi, a_lo + 2 * k16), ptx::smem_desc_join(hi, b_lo + 2 * k16),
                             idesc, acc);
            if (split)
              ptx::mma_bf16_ss(tmem_d1, ptx::smem_desc_join(hi, a_lo + 2 * k16),
                               ptx::smem_desc_join(hi, b_lo + b2_off + 2 * k16), idesc, acc);
          }
        }
        CSD_TSM(2);
        ptx::mma_commit(ebar);  // frees the stage when the MMAs above have read it
        CSD_TSM(3);
        }
        __syncwarp();
        accumulate = 1u;
        stage = nstage;
        par = npar;
        a_lo = n_lo;
      }
      if (ptx::elect_one()) ptx::mma_commit(tmem_full_bar);       // accumulator complete
      __syncwarp();
    }
  } else {
    // ===== epilogue (warps 2..5): warp q owns TMEM lanes [32q, 32q+32) with q = warp % 4 =====
    const int q = warp & 3;
    const int m = q * 32 + lane;  // row of the tile
    const int wl = m % p.TW;
    const int hl = (m / p.TW) % p.TH;
    const int bl = m / (p.TW * p.TH);
    const int b = b0 + bl, h = h0 + hl, w = w0 + wl;
    const bool valid = (bl < p.TB) && (b < p.B) && (h < p.H) && (w < p.W);
    const long long pix = ((long long)b * p.H + h) * p.W + w;

    const float* s_add = build_addend_table(p, s_add_addr, b0, n0, threadIdx.x - 64);
    if (s_add != nullptr && valid) s_add += bl * p.n_tile;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tcgen05_fence_after();
    if (threadIdx.x == 64) CSD_TS(5);
    epilogue_rows<kTf32>(p, tmem_base + ((uint32_t)(q * 32) << 16), valid, pix, b, n0, z, s_add);
    if (threadIdx.x == 64) CSD_TS(6);
  }

  // teardown: everyone is done with TMEM before the allocating warp frees it
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (lane == 0) CSD_TS(7);
  }
}


// ---------------------------------------------------------------------------------------------------
// Halo mode (3x3 stride-1 convolutions at the large resolutions).
//
// Per 32-channel chunk ONE TMA box load brings the (16*mt+2) x 10 pixel halo of the CTA's mt stacked
// 16x8-pixel tiles into shared memory (64-byte swizzled rows, one row per pixel). All 9 taps of every
// tile are then issued from that single copy: the UMMA descriptor of tap (ky,kx) of tile t starts
// ((16t+ky)*10 + kx) rows into the halo and strides 10 rows between 8-row core-matrix groups (SBO).
// This works because the hardware applies the 64-byte swizzle to absolute shared-memory address
// bits, so a start address shifted by whole 64-byte rows stays consistent with what TMA wrote
// (verified on B200 with tools/umma_halo_probe.cu). Compared with the per-tap kernel above the
// activation bytes moved from L2 drop 9x -> ~1.3x of the tile, and with mt = 2 every weight slab
// feeds two tiles. Weights stream through their own ring, one N x 64-byte slab per (chunk, tap).
// ---------------------------------------------------------------------------------------------------
constexpr int kHaloTW = 8, kHaloTH = 16;
constexpr int kMaxAStages = 4, kMaxBStages = 8;

__global__ void __launch_bounds__(kConvThreads)
conv_halo_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                 const __grid_constant__ CUtensorMap mapB, const ConvGemmKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + p.a_stages * p.a_stage_bytes;
  const uint32_t bar_base = b_base + p.b_stages * p.b_stage_bytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kMaxAStages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * kMaxAStages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * kMaxAStages + kMaxBStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kMaxAStages + 2 * kMaxBStages);
  const uint32_t tmem_slot = tmem_full_bar + 8u;
  const uint32_t s_add_addr = tmem_full_bar + 16u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x;
  const int tw = t % p.tiles_w;
  const int th = (t / p.tiles_w) % p.tiles_h;
  const int b = t / (p.tiles_w * p.tiles_h);
  const int w0 = tw * kHaloTW, h0 = th * kHaloTH * p.mt;
  const int n0 = blockIdx.y * p.n_tile;
  if (threadIdx.x == 0) CSD_TS(0);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapA0);
    ptx::prefetch_tensormap(&mapB);
    for (int s = 0; s < p.a_stages; ++s) { ptx::mbar_init(a_full(s), 1); ptx::mbar_init(a_empty(s), 1); }
    for (int s = 0; s < p.b_stages; ++s) { ptx::mbar_init(b_full(s), 1); ptx::mbar_init(b_empty(s), 1); }
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

  if (threadIdx.x == 0) CSD_TS(1);
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int ia = 0, ib = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const CUtensorMap* mapA = (s == 0) ? &mapA0 : (s == 1) ? &mapA1 : (s == 2) ? &mapA2 : &mapA3;
        const int taps = p.seg_taps[s];
        const int halo = (taps == 9) ? 1 : 0;
        const uint32_t a_bytes = (uint32_t)((kHaloTW + 2 * halo) * (kHaloTH * p.mt + 2 * halo) * kRowBytes);
        const int nchunks = p.seg_chunks[s];
        for (int c = 0; c < nchunks; ++c, ++ia) {
          const int sa = ia % p.a_stages;
          if (!((p.debug_nodata & 2) && ia >= p.a_stages)) {
          ptx::mbar_wait(a_empty(sa), ((ia / p.a_stages) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(a_full(sa), a_bytes);
          ptx::tma_load_4d(a_base + sa * p.a_stage_bytes, mapA, a_full(sa), p.seg_coff[s] + c * kChunkK, w0 - halo,
                           h0 - halo, b);
          }
          for (int tap = 0; tap < taps; ++tap, ++ib) {
            const int sb = ib % p.b_stages;
            if ((p.debug_nodata & 1) && ib >= p.b_stages) continue;
            ptx::mbar_wait(b_empty(sb), ((ib / p.b_stages) & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(b_full(sb), p.b_box_bytes * p.nsplit);
            const int kidx = p.seg_kbase[s] + tap * nchunks + c;
            for (int j = 0; j < p.nsplit; ++j)
              ptx::tma_load_3d(b_base + sb * p.b_stage_bytes + j * p.b_box_bytes, &mapB, b_full(sb),
                               p.wt_k_off + kidx * kChunkK, n0 + j * p.n_sub, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16_m128((uint32_t)p.n_sub);
      int ia = 0, ib = 0;
      uint32_t accumulate = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const int taps = p.seg_taps[s];
        const int halo = (taps == 9) ? 1 : 0;
        const int pitch = kHaloTW + 2 * halo;              // pixels per halo row
        const uint32_t sbo = (uint32_t)pitch * kRowBytes;  // stride between 8-pixel core-matrix groups
        for (int c = 0; c < p.seg_chunks[s]; ++c, ++ia) {
          const int sa = ia % p.a_stages;
          if (!((p.debug_nodata & 2) && ia >= p.a_stages)) ptx::mbar_wait(a_full(sa), (ia / p.a_stages) & 1);
          ptx::tcgen05_fence_after();
          if (ia == 0) CSD_TS(3);
          if (ia == 1) CSD_TS(8);
          const uint32_t a_addr = a_base + sa * p.a_stage_bytes;
          for (int tap = 0; tap < taps; ++tap, ++ib) {
            const int sb = ib % p.b_stages;
            if (!((p.debug_nodata & 1) && ib >= p.b_stages)) ptx::mbar_wait(b_full(sb), (ib / p.b_stages) & 1);
            ptx::tcgen05_fence_after();
            const uint32_t b_addr = b_base + sb * p.b_stage_bytes;
            const int dy = halo ? tap / 3 : 0, dx = halo ? tap % 3 : 0;
            for (int tt = 0; tt < p.mt; ++tt) {
              const uint32_t a_tile = a_addr + (uint32_t)(((kHaloTH * tt + dy) * pitch + dx) * kRowBytes);
#pragma unroll
              for (int kk = 0; kk < kChunkK / 16; ++kk) {
                const uint64_t a_desc = ptx::make_smem_desc(a_tile + kk * 32, 16, sbo, kLayoutSw64);
                for (int j = 0; j < p.nsplit; ++j) {
                  const uint64_t b_desc =
                      ptx::make_smem_desc(b_addr + j * p.b_box_bytes + kk * 32, 16, 512, kLayoutSw64);
                  ptx::mma_bf16_ss(tmem_base + tt * p.n_tile + j * p.n_sub, a_desc, b_desc, idesc,
                                   (accumulate || kk > 0) ? 1u : 0u);
                }
              }
            }
            accumulate = 1;
            ptx::mma_commit(b_empty(sb));
          }
          ptx::mma_commit(a_empty(sa));
        }
      }
      CSD_TS(4);
      ptx::mma_commit(tmem_full_bar);
    }
  } else {
    // ===== epilogue: warp q owns TMEM lanes [32q, 32q+32); one pass per stacked tile =====
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const float* s_add = build_addend_table(p, s_add_addr, b, n0, threadIdx.x - 64);
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tcgen05_fence_after();
    if (threadIdx.x == 64) CSD_TS(5);
    for (int tt = 0; tt < p.mt; ++tt) {
      const int h = h0 + kHaloTH * tt + m / kHaloTW, w = w0 + m % kHaloTW;
      const bool valid = (h < p.H) && (w < p.W);
      const long long pix = ((long long)b * p.H + h) * p.W + w;
      epilogue_rows<false>(p, tmem_base + ((uint32_t)(q * 32) << 16) + tt * p.n_tile, valid, pix, b, n0, 0, s_add);
    }
    if (threadIdx.x == 64) CSD_TS(6);
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (lane == 0) CSD_TS(7);
  }
}

// ---------------------------------------------------------------------------------------------------
// Transposed halo mode ("weights are the M operand") for 3x3 stride-1 convolutions.
//
// tcgen05.mma with cta_group::1 costs ~100 cycles per M=128 instruction however small N is (measured,
// tools/mma_rate_probe.cu: N=96 -> 101 cycles, N=256 -> 128 cycles), so with pixels on M a 96-channel
// layer can use at most 47% of the tensor pipe. Here the roles are swapped: D^T[c_out, pixel] with
// M = 128 output channels (rows of the weight slab, zero rows beyond C_out come from TMA's
// out-of-bounds fill) and N = 256 pixels (a 32 x 8 pixel macro-tile). The pixel operand is the same
// single halo buffer as in conv_halo_kernel - (32+2) x 10 pixels, one TMA box per 32-channel chunk -
// addressed per tap by a shifted descriptor (SBO = 10 rows). One instruction then does 128 x 256 x 16
// MACs in 128 cycles.
// Epilogue: TMEM lane = output channel, column = pixel; bias / temb are per-lane scalars, the
// per-channel GroupNorm partial sums of the stored output are free per-thread accumulations.
//
// Fused GroupNorm + SiLU prologue (segments with a `norm` table): dedicated warps transform every pixel-halo
// stage in place between TMA arrival and MMA issue:
// y = SiLU(x * scale[b,c] + shift[b,c]) on the 16-byte units of the swizzled tile (physical unit j' of row
// r holds logical channels 8*(j' ^ ((r >> 1) & 3)) .. +8 under SWIZZLE_64B), skipping halo pixels outside
// the image (they must stay the zeros TMA wrote: the reference pads the *normalised* activation,
// models/layers.py:119-132 padding=1 after models/layerspp.py:242). A stage then goes
// TMA -> a_full -> transform -> fence.proxy.async -> a_ready -> MMA -> a_empty. One halo stage is transformed
// once and feeds all 9 taps, so the MUFU work is 1/9 of what a per-tap operand transform would cost.
// ---------------------------------------------------------------------------------------------------
constexpr int kTPix = 256;      // TMEM columns per accumulator = max pixels per CTA tile (N of the MMA)
constexpr int kTRows = 32;      // max image rows per macro tile (tiles are t_rows x 8 pixels, t_rows = 32 or 20)
constexpr int kTChan = 128;     // output channels per CTA (M of the MMA)

// ---------------------------------------------------------------------------------------------------
// The transposed halo kernel (mode 2), persistent.
//
// A one-tile-per-CTA version of this kernel spent ~45% of each CTA's life outside its main loop (pipeline fill ~3.5k
// cycles, epilogue ~8.7k cycles against ~6.9k tensor cycles for a 96-channel layer; measured with
// tools/conv_phase_timing.py) and relied on a second resident CTA to fill the gap. Here ONE CTA per SM
// walks a static list of tiles (tile = blockIdx.x + i * gridDim.x) with five specialised roles:
//   warp 0      TMA producer A    pixel halos (own thread: runs a whole tile ahead of the weights)
//   warps 18-20 TMA producers B   weight slabs, round-robin: one thread starts a TMA load only every ~360 cycles
//                                 (tools/tma_rate_probe.cu) and a chunk needs nine 8 KB slabs per 2304 tensor cycles
//   warp 1      MMA issuer        accumulator alternates between two 256-column TMEM buffers
//   warps 2-9   operand transform fused GroupNorm(+SiLU) of every pixel-halo stage (when a segment has `norm`)
//   warps 10-17 epilogue          TMEM -> (scale, bias, temb, GN partial sums) -> staging smem -> one TMA store
// so the epilogue of tile i, the main loop of tile i+1 and the loads/transforms of tile i+2 overlap, and
// the fill/drain cost is paid once per SM instead of once per tile. The whole 227 KB of shared memory
// belongs to the CTA: 3 pixel-halo buffers, up to 12 weight slabs, a dedicated 64 KB staging tile.
// ---------------------------------------------------------------------------------------------------
#ifndef CSD_PB_PRODUCERS
#define CSD_PB_PRODUCERS 3
#endif
constexpr int kPBProducers = CSD_PB_PRODUCERS;                     // weight-slab producer warps (issue-rate bound)
constexpr int kPThreads = 576 + 32 * kPBProducers;
constexpr int kPTransformThreads = 256;
constexpr int kPEpiThreads = 256;
constexpr int kPMaxBStages = 12;
constexpr int kPStagePitch = 128;                                  // staging row pitch (channels), constant
constexpr int kPStagingBytes = kTPix * kPStagePitch * 2;           // 64 KB
constexpr int kPBarBytes = 8 * (3 * kMaxAStages + 2 * kPMaxBStages + 4 + 1) + 8;   // barriers + tmem slot, 16-aligned

struct TileCoord {
  int b, h0, w0, n0, sp;
};
// i-th (segment, chunk) of a tile in schedule order
__device__ __forceinline__ void sched_at(const ConvGemmKernelParams& p, int i, int& s, int& c) {
  if (p.sched_tab) {
    s = p.sched_seg[i];
    c = p.sched_chunk[i];
  } else {
    s = 0;
    while (i >= p.seg_chunks[s]) { i -= p.seg_chunks[s]; ++s; }
    c = i;
  }
}
__device__ __forceinline__ TileCoord decode_tile(const ConvGemmKernelParams& p, int tile) {
  TileCoord t;
  const int nb = tile % p.n_blocks;
  t.sp = tile / p.n_blocks;
  const int tw = t.sp % p.tiles_w;
  const int r = t.sp / p.tiles_w;
  const int th = r % p.tiles_h;
  t.b = r / p.tiles_h;
  t.w0 = tw * kHaloTW;
  t.h0 = th * p.t_rows;
  t.n0 = nb * kTChan;
  return t;
}

template <bool kTf32>
__global__ void __launch_bounds__(kPThreads, 1)
conv_halo_tp_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                    const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapOut,
                    const __grid_constant__ CUtensorMap mapOut2, const __grid_constant__ CUtensorMap mapW0,
                    const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                    const __grid_constant__ CUtensorMap mapW3, const ConvGemmKernelParams p) {
  // kTf32: the reference-precision plan. Activations, weights and the output are fp32 words, the MMA is kind::tf32.
  // Every shared-memory structure keeps its BYTE geometry (64-byte rows, SWIZZLE_64B, 2 MMAs of 32 bytes of K per
  // (chunk, tap)): a chunk is 16 fp32 channels instead of 32 bf16 channels, so only the chunk counts double.
  constexpr int CH = kTf32 ? 16 : kChunkK;                             // channels per 64-byte chunk
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;                                   // pixel halos
  const uint32_t b_base = a_base + p.a_stages * p.a_stage_bytes;       // weight slabs
  const uint32_t stage_base = b_base + p.b_stages * p.b_stage_bytes;   // epilogue staging tile (1 KB aligned)
  const uint32_t bar_base = stage_base + (uint32_t)p.staging_bytes;
  const uint32_t a_full0 = bar_base, a_empty0 = a_full0 + 8u * kMaxAStages, a_ready0 = a_empty0 + 8u * kMaxAStages;
  const uint32_t b_full0 = a_ready0 + 8u * kMaxAStages, b_empty0 = b_full0 + 8u * kPMaxBStages;
  const uint32_t tmem_full0 = b_empty0 + 8u * kPMaxBStages, tmem_empty0 = tmem_full0 + 16u;
  const uint32_t tmem_slot = tmem_empty0 + 16u;
  const uint32_t coef_addr = bar_base + kPBarBytes;                    // float2 per padded K channel

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapA0);
    ptx::prefetch_tensormap(&mapB);
    ptx::prefetch_tensormap(&mapOut);
    for (int s = 0; s < p.a_stages; ++s) {
      ptx::mbar_init(a_full0 + 8u * s, 1);
      ptx::mbar_init(a_empty0 + 8u * s, 1);
      ptx::mbar_init(a_ready0 + 8u * s, kPTransformThreads);
    }
    for (int s = 0; s < p.b_stages; ++s) { ptx::mbar_init(b_full0 + 8u * s, 1); ptx::mbar_init(b_empty0 + 8u * s, 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(tmem_full0 + 8u * s, 1); ptx::mbar_init(tmem_empty0 + 8u * s, kPEpiThreads); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * kTPix);   // both accumulators: all 512 columns (one CTA per SM)
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===== TMA producer A: pixel halos. Separate from the weight producer so that the halo ring can run a whole
    //       tile ahead (the transform needs that lead) instead of being throttled by the weight ring's depth =====
    if (lane == 0) {
      uint32_t sa = 0, a_par = 1;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        CSD_TSP(10);
        for (int i = 0; i < p.n_sched; ++i) {
          int s, c;
          sched_at(p, i, s, c);
          const CUtensorMap* mapA = (s == 0) ? &mapA0 : (s == 1) ? &mapA1 : (s == 2) ? &mapA2 : &mapA3;
          const int halo = (p.seg_taps[s] == 9) ? 1 : 0;
          const uint32_t a_bytes = (uint32_t)((kHaloTW + 2 * halo) * (p.t_rows + 2 * halo) * kRowBytes);
          if (p.debug_nodata & 2) continue;
          ptx::mbar_wait(a_empty0 + 8u * sa, a_par);
          if (i == 0) CSD_TSP(11);
          ptx::mbar_arrive_expect_tx(a_full0 + 8u * sa, a_bytes);
          ptx::tma_load_4d(a_base + sa * p.a_stage_bytes, mapA, a_full0 + 8u * sa, p.seg_coff[s] + c * CH,
                           tc.w0 - halo, tc.h0 - halo, tc.b);
          if (++sa == (uint32_t)p.a_stages) { sa = 0; a_par ^= 1u; }
        }
      }
    }
  } else if (warp >= 18) {
    // ===== TMA producer B: weights. A ring slot holds THREE slabs = the three taps of one kernel row for one channel
    //       chunk, filled by ONE TMA op through the segment's 3-D map (k, row, tap) - the tap dimension has the smaller
    //       stride, tools/tma_taps_probe.cu - so barrier round trips, TMA ops and commits are paid once per six MMAs.
    //       (What the weight stream costs is per transaction, not per byte: profiles/conv_nodata_r2.txt.) A 1-tap chunk
    //       takes one slot with a single slab through the 2-D map. One producer thread is enough now (an op every
    //       ~360 cycles against one needed every 6 MMAs); warps 19-20 idle. =====
    if (lane == 0 && warp == 18) {
      uint32_t sb = 0, b_par = 1;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        for (int i = 0; i < p.n_sched; ++i) {
          int s, c;
          sched_at(p, i, s, c);
          if (p.debug_nodata & 1) continue;   // perf experiment: no weight traffic
          if (p.seg_taps[s] == 9) {
            const CUtensorMap* mapW = (s == 0) ? &mapW0 : (s == 1) ? &mapW1 : (s == 2) ? &mapW2 : &mapW3;
            for (int g3 = 0; g3 < 3; ++g3) {
              ptx::mbar_wait(b_empty0 + 8u * sb, b_par);
              ptx::mbar_arrive_expect_tx(b_full0 + 8u * sb, 3u * kTChan * kRowBytes);
              ptx::tma_load_3d(b_base + sb * p.b_stage_bytes, mapW, b_full0 + 8u * sb, c * CH, tc.n0, 3 * g3);
              if (++sb == (uint32_t)p.b_stages) { sb = 0; b_par ^= 1u; }
            }
          } else {
            const int kcol = p.wt_k_off + p.seg_kbase[s] * kChunkK + c * CH;
            ptx::mbar_wait(b_empty0 + 8u * sb, b_par);
            ptx::mbar_arrive_expect_tx(b_full0 + 8u * sb, (uint32_t)kTChan * kRowBytes);
            ptx::tma_load_3d(b_base + sb * p.b_stage_bytes, &mapB, b_full0 + 8u * sb, kcol, tc.n0, 0);
            if (++sb == (uint32_t)p.b_stages) { sb = 0; b_par ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // The whole warp walks the loop converged and one elected lane issues: ring indices, parities and descriptors are
    // then warp-uniform values the compiler keeps in the uniform datapath. With the loop inside `if (lane == 0)` every
    // tcgen05 operand went through R2UR moves from a divergent thread's registers, and the issue loop - not the tensor
    // pipe - set the pace: ~168 cycles per MMA whatever N (profiles/conv_phase_timing_r2.txt: N = 160 tiles at 40 px
    // cost as much per instruction as N = 256 tiles, with or without operand traffic), against 80 / 112 / 128 cycles for
    // N = 160 / 224 / 256 in a bare issue loop (tools/mma_rate_probe3.cu).
    {
      const uint32_t idesc = kTf32 ? ptx::make_idesc_tf32_m128((uint32_t)p.t_pix) : ptx::make_idesc_bf16_m128((uint32_t)p.t_pix);
      const uint32_t w_hi = ptx::smem_desc_hi(512, kLayoutSw64);
      auto mma = [](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t acc) {
        if (kTf32) ptx::mma_tf32_ss(d, ad, bd, id, acc);
        else ptx::mma_bf16_ss(d, ad, bd, id, acc);
      };
      const uint32_t a_go0 = p.has_norm ? a_ready0 : a_full0;
      // Everything the inner loop touches lives in registers and advances by additions: ring slot -> (descriptor low
      // word, full / empty barrier address). One thread issues in order, so every dependent load or multiply in front of
      // an MMA is exposed latency: with `b_base + sb * p.b_stage_bytes` (a constant-bank load + IMAD + R2UR chain per tap)
      // the bare loop ran at ~148 cycles per MMA whatever N (profiles/conv_phase_timing_r2.txt), the tensor pipe's own
      // rate being N/2 = 80 / 112 / 128 cycles (tools/mma_rate_probe3.cu).
      const uint32_t n_b = (uint32_t)p.b_stages, n_a = (uint32_t)p.a_stages;
      const uint32_t w_step = p.b_stage_bytes >> 4, x_step = p.a_stage_bytes >> 4;
      const uint32_t w_lo_first = ptx::smem_desc_lo(b_base, 16), x_lo_first = ptx::smem_desc_lo(a_base, 16);
      const int n_sched = p.n_sched;
      const uint32_t nodata = (uint32_t)p.debug_nodata;
      uint32_t sa = 0, a_par = 0, sb = 0, b_par = 0;
      uint32_t w_lo = w_lo_first, bf_bar = b_full0, be_bar = b_empty0;      // weight ring cursor
      uint32_t x_lo_s = x_lo_first, ag_bar = a_go0, ae_bar = a_empty0;      // halo ring cursor
      uint32_t acc = 0, acc_par = 1;   // a fresh tmem_empty barrier passes a wait on parity 1
      constexpr int pitch9 = kHaloTW + 2;
      const uint32_t x_hi9 = ptx::smem_desc_hi(pitch9 * kRowBytes, kLayoutSw64);
      const uint32_t x_hi1 = ptx::smem_desc_hi(kHaloTW * kRowBytes, kLayoutSw64);
      const uint64_t w_desc_hi = static_cast<uint64_t>(w_hi) << 32;
      const uint64_t x9_desc_hi = static_cast<uint64_t>(x_hi9) << 32, x1_desc_hi = static_cast<uint64_t>(x_hi1) << 32;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        if (lane == 0) CSD_TSP(12);
        ptx::mbar_wait(tmem_empty0 + 8u * acc, acc_par);   // epilogue has drained this accumulator
        ptx::tcgen05_fence_after();
        if (lane == 0) CSD_TSP(0);
        const uint32_t d_tmem = tmem_base + acc * kTPix;
        uint32_t accumulate = 0;
        for (int i = 0; i < n_sched; ++i) {
          int s, c;
          sched_at(p, i, s, c);
          const bool nine = p.seg_taps[s] == 9;
          if (!(nodata & 2)) ptx::mbar_wait(ag_bar, a_par);
          ptx::tcgen05_fence_after();
          if (lane == 0 && i == 0) CSD_TSP(1);
          if (lane == 0 && i == 1) CSD_TSP(13);
          if (nine) {
            // one ring slot = the three slabs of a kernel row: one wait, six MMAs, one commit
#pragma unroll
            for (int g3 = 0; g3 < 3; ++g3) {
              if (!(nodata & 1)) ptx::mbar_wait(bf_bar, b_par);
              ptx::tcgen05_fence_after();
              const uint32_t x_row = x_lo_s + (uint32_t)((g3 * pitch9 * kRowBytes) >> 4);
              if (ptx::elect_one()) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                  const uint32_t wl = w_lo + (uint32_t)((j * kTChan * kRowBytes) >> 4);
                  const uint32_t x_lo = x_row + (uint32_t)((j * kRowBytes) >> 4);
                  mma(d_tmem, w_desc_hi | wl, x9_desc_hi | x_lo, idesc, (g3 == 0 && j == 0) ? accumulate : 1u);
                  mma(d_tmem, w_desc_hi | (wl + 2), x9_desc_hi | (x_lo + 2), idesc, 1u);
                }
                ptx::mma_commit(be_bar);
                if (g3 == 2) ptx::mma_commit(ae_bar);
              }
              __syncwarp();
              const bool wrap = sb + 1 == n_b;
              sb = wrap ? 0u : sb + 1;
              b_par = wrap ? b_par ^ 1u : b_par;
              w_lo = wrap ? w_lo_first : w_lo + w_step;
              bf_bar = wrap ? b_full0 : bf_bar + 8u;
              be_bar = wrap ? b_empty0 : be_bar + 8u;
            }
            accumulate = 1u;
          } else {
            if (!(nodata & 1)) ptx::mbar_wait(bf_bar, b_par);
            ptx::tcgen05_fence_after();
            if (ptx::elect_one()) {
              mma(d_tmem, w_desc_hi | w_lo, x1_desc_hi | x_lo_s, idesc, accumulate);
              mma(d_tmem, w_desc_hi | (w_lo + 2), x1_desc_hi | (x_lo_s + 2), idesc, 1u);
              ptx::mma_commit(be_bar);
              ptx::mma_commit(ae_bar);
            }
            __syncwarp();
            accumulate = 1u;
            const bool wrap = sb + 1 == n_b;
            sb = wrap ? 0u : sb + 1;
            b_par = wrap ? b_par ^ 1u : b_par;
            w_lo = wrap ? w_lo_first : w_lo + w_step;
            bf_bar = wrap ? b_full0 : bf_bar + 8u;
            be_bar = wrap ? b_empty0 : be_bar + 8u;
          }
          const bool awrap = sa + 1 == n_a;
          sa = awrap ? 0u : sa + 1;
          a_par = awrap ? a_par ^ 1u : a_par;
          x_lo_s = awrap ? x_lo_first : x_lo_s + x_step;
          ag_bar = awrap ? a_go0 : ag_bar + 8u;
          ae_bar = awrap ? a_empty0 : ae_bar + 8u;
        }
        if (lane == 0) CSD_TSP(2);
        if (ptx::elect_one()) ptx::mma_commit(tmem_full0 + 8u * acc);
        __syncwarp();
        if (++acc == 2u) { acc = 0; acc_par ^= 1u; }
      }
    }
  } else if (warp < 10) {
    // ===== operand transform: fused GroupNorm(+SiLU) of the pixel halos =====
    // Thread (j, rg): j = logical 16-byte unit (8 channels) of the 64-byte rows, rg = row group. Its 8
    // (scale, shift) pairs sit in registers for the whole stage; it walks rows rg, rg+32, ... two at a time
    // (independent load -> math -> store chains; the stage is shared-memory-latency bound, hence 8 warps).
    // A warp touches 8 consecutive rows x 4 units = 512 contiguous
    // bytes per access: conflict free. y = SiLU(x*sc + sh) = h + h*tanh(h), h = x*(sc/2) + sh/2: 3 FP + 1 MUFU.
    if (p.has_norm) {
      const int tt = threadIdx.x - 64;   // 0..255
      const int j = tt & 3, rg = tt >> 2;
      float2* tab = reinterpret_cast<float2*>(__cvta_shared_to_generic(coef_addr));
      uint32_t sa = 0, a_par = 0;
      int tab_b = -1;
      // The tiles of a CTA are gridDim.x apart - more than the tiles of one image - so EVERY tile starts with a new
      // image's (scale, shift) table, and these warps sit within a few hundred cycles per tile of the kernel's critical
      // path (an experiment that added two dependent phases to the table build, +0.34 us per tile, slowed the whole PC
      // step by exactly that: profiles/inline_groupnorm_experiment_r2.txt). The table's global loads (an L2 round trip:
      // every image is new to this SM) are therefore issued one tile AHEAD, into registers, and only written to shared
      // memory at the tile boundary. Up to kPre table slots per thread; wider tables keep the load-at-the-boundary path.
      constexpr int kPre = 2;
      int tab_total = 0;
      for (int s = 0; s < p.nseg; ++s) tab_total += p.seg_chunks[s] * CH;
      const bool prefetch = tab_total <= kPre * kPTransformThreads && !(p.debug_nodata & 64);
      const float2* psrc[kPre];
      int pcnt[kPre];
      float ppre[kPre];
      float2 nv[kPre];
#pragma unroll
      for (int q = 0; q < kPre; ++q) {
        psrc[q] = nullptr;
        pcnt[q] = 0;
        ppre[q] = 1.0f;
        nv[q] = make_float2(0.f, 0.f);
        const int idx = tt + q * kPTransformThreads;
        int base = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int nch = p.seg_chunks[s] * CH;
          if (idx >= base && idx < base + nch) {
            const int i = idx - base;
            if (p.seg_norm[s] != nullptr && i < p.seg_ccnt[s]) {
              psrc[q] = reinterpret_cast<const float2*>(p.seg_norm[s]) + i;
              pcnt[q] = p.seg_ccnt[s];
            }
            ppre[q] = (p.seg_silu[s] && !kTf32) ? 0.5f : 1.0f;
          }
          base += nch;
        }
      }
      auto fetch = [&](int b_img) {
#pragma unroll
        for (int q = 0; q < kPre; ++q)
          nv[q] = psrc[q] != nullptr ? __ldg(psrc[q] + (long long)b_img * pcnt[q]) : make_float2(0.f, 0.f);
      };
      if (prefetch && (int)blockIdx.x < num_tiles) fetch(decode_tile(p, blockIdx.x).b);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        if (tc.b != tab_b && prefetch) {
          asm volatile("bar.sync 2, 256;" ::: "memory");     // nobody still reads the previous image's table
#pragma unroll
          for (int q = 0; q < kPre; ++q) {
            const int idx = tt + q * kPTransformThreads;
            if (idx < tab_total) tab[idx] = make_float2(nv[q].x * ppre[q], nv[q].y * ppre[q]);
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
          tab_b = tc.b;
        }
        if (prefetch && tile + (int)gridDim.x < num_tiles) fetch(decode_tile(p, tile + gridDim.x).b);
        if (tc.b != tab_b) {             // (scale, shift) of this image's channels, all segments back to back
          asm volatile("bar.sync 2, 256;" ::: "memory");
          int base = 0;
          for (int s = 0; s < p.nseg; ++s) {
            const int nch = p.seg_chunks[s] * CH;
            const float2* src = reinterpret_cast<const float2*>(p.seg_norm[s]);
            // bf16 SiLU path stores (sc/2, sh/2) for y = h + h tanh(h); the fp32 path evaluates x sigmoid(x) directly
            const float pre = (p.seg_silu[s] && !kTf32) ? 0.5f : 1.0f;
            for (int i = tt; i < nch; i += kPTransformThreads) {
              float2 v = make_float2(0.f, 0.f);
              if (src != nullptr && i < p.seg_ccnt[s]) v = __ldg(src + (long long)tc.b * p.seg_ccnt[s] + i);
              tab[base + i] = make_float2(v.x * pre, v.y * pre);
            }
            base += nch;
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
          tab_b = tc.b;
        }
        for (int i = 0; i < p.n_sched; ++i) {
          int s, c;
          sched_at(p, i, s, c);
          const int base = p.seg_tab_base[s];
          const bool norm = p.seg_norm[s] != nullptr;
          const bool act = p.seg_silu[s] != 0;
          const int halo = (p.seg_taps[s] == 9) ? 1 : 0;
          const int pitch = kHaloTW + 2 * halo;
          const int rows = pitch * (p.t_rows + 2 * halo);
          {
            ptx::mbar_wait(a_full0 + 8u * sa, a_par);
            if (tt == 0 && i == 0) CSD_TSP(8);
            if (norm) {
              uint4* st = reinterpret_cast<uint4*>(__cvta_shared_to_generic(a_base + sa * p.a_stage_bytes));
              // (scale, shift) pairs of this thread's 16-byte unit: 8 bf16 channels (4 float4) or 4 fp32 channels (2)
              constexpr int NCF = kTf32 ? 2 : 4;
              float4 cf[4];
#pragma unroll
              for (int i = 0; i < NCF; ++i)
                cf[i] = *reinterpret_cast<const float4*>(tab + base + c * CH + j * (2 * NCF) + 2 * i);
              auto row_unit = [&](int r) -> int {     // 16-byte unit index of this thread's channels in row r, or -1
                const int hy = halo ? r / (kHaloTW + 2) : (r >> 3);
                const int hx = r - hy * pitch;
                const int gh = tc.h0 - halo + hy, gw = tc.w0 - halo + hx;
                const bool ok = r < rows && gh >= 0 && gh < p.H && gw >= 0 && gw < p.W;   // conv padding stays zero
                return ok ? r * 4 + (j ^ ((r >> 1) & 3)) : -1;                          // SWIZZLE_64B unit
              };
              auto apply = [&](uint4 raw) -> uint4 {
                if (kTf32) {      // 4 fp32 channels; the result feeds the tensor core only: rounded to tf32 (nearest)
                  float f[4] = {__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z),
                                __uint_as_float(raw.w)};
#pragma unroll
                  for (int i = 0; i < 2; ++i) {
                    float y0 = fmaf(f[2 * i], cf[i].x, cf[i].y), y1 = fmaf(f[2 * i + 1], cf[i].z, cf[i].w);
                    if (act) {    // tanh.approx would cost 2^-11 (the tf32 rounding itself); ex2.approx + rcp.approx cost
                                  // ~2^-21 and a third of the instructions of an IEEE division - the transform warps were
                                  // the tf32 main loop's limiter (4.9k cycles per chunk against 2.3k of tensor work)
                      y0 = __fdividef(y0, 1.f + __expf(-y0));
                      y1 = __fdividef(y1, 1.f + __expf(-y1));
                    }
                    f[2 * i] = round_tf32(y0);
                    f[2 * i + 1] = round_tf32(y1);
                  }
                  return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                                    __float_as_uint(f[3]));
                }
                bf16x8 v;
                *reinterpret_cast<uint4*>(&v) = raw;
                float f[8];
                unpack8(v, f);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float y0 = fmaf(f[2 * i], cf[i].x, cf[i].y), y1 = fmaf(f[2 * i + 1], cf[i].z, cf[i].w);
                  if (act) {
                    float t0, t1;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(y0));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(y1));
                    y0 = fmaf(y0, t0, y0);
                    y1 = fmaf(y1, t1, y1);
                  }
                  f[2 * i] = y0;
                  f[2 * i + 1] = y1;
                }
                v = pack8(f);
                return *reinterpret_cast<const uint4*>(&v);
              };
              for (int r = rg; r < rows; r += 128) {
                const int u0 = row_unit(r), u1 = row_unit(r + 64);
                uint4 x0 = make_uint4(0, 0, 0, 0), x1 = x0;
                if (u0 >= 0) x0 = st[u0];
                if (u1 >= 0) x1 = st[u1];
                x0 = apply(x0);
                x1 = apply(x1);
                if (u0 >= 0) st[u0] = x0;
                if (u1 >= 0) st[u1] = x1;
              }
              ptx::fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async reads
            }
            ptx::mbar_arrive(a_ready0 + 8u * sa);
            if (tt == 0 && i == 0) CSD_TSP(9);
            if (++sa == (uint32_t)p.a_stages) { sa = 0; a_par ^= 1u; }
          }
        }
      }
    }
  } else if (kTf32) {
    // ===== epilogue, fp32 output (8 warps) =====
    // Same thread <-> channel mapping as the bf16 epilogue below (thread = output channel, warps 10-13 the first half of
    // the tile's pixel rows, warps 14-17 the second). An fp32 [256 pixels x 128 channels] staging tile would be 128 KB,
    // so the tile leaves in two passes over the pixel rows: pass A = the first 8 rows of each half, pass B = the rest
    // (8 / 6 / 4 / 2 rows for 32 / 28 / 24 / 20-row tiles), each through a [2 halves x 64 pixels x 128 channels] fp32
    // staging tile (64 KB) and one TMA store per half (mapOut: 8-row boxes, mapOut2: the remainder rows). All 256
    // threads work in both passes. The residual of the fp32 plan is added here from global memory (an identity K
    // segment would be truncated to tf32 by the tensor core and bias the residual stream of every block towards zero).
    const int et = threadIdx.x - 320;
    const int q = warp & 3;
    const int half = (warp - 10) >> 2;
    const int cl = q * 32 + lane;
    const int spitch = p.out_box_c;                   // staging row pitch = channels per store box
    const int half_pix = p.t_pix >> 1, half_rows = p.t_rows >> 1;
    float* stage_h = reinterpret_cast<float*>(__cvta_shared_to_generic(stage_base)) + half * (64 * spitch) + cl;
    const uint32_t stage_h_addr = stage_base + (uint32_t)(half * 64 * spitch * 4);
    const bool has_stats = p.stat_partials != nullptr;
    const float scale = p.scale;
    const float* res = reinterpret_cast<const float*>(p.res);
    uint32_t acc = 0, full_par = 0;
    bool store_pending = false;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile);
      const int c = tc.n0 + cl;
      const int cb = min(kTChan, p.n_store - tc.n0);
      const bool c_valid = cl < cb;
      const float add_cs = c_valid ? scale * ((p.bias != nullptr ? __ldg(p.bias + c) : 0.f) +
                                              (p.temb != nullptr ? __ldg(p.temb + (long long)tc.b * p.temb_pitch + c) : 0.f))
                                   : 0.f;
      const int w_lim = min(kHaloTW, p.W - tc.w0);
      const int m_lim_h = (w_lim == kHaloTW) ? min(half_pix, min(p.t_pix, (p.H - tc.h0) * kHaloTW) - half * half_pix) : 0;
      if (et == 0) CSD_TSP(3);
      ptx::mbar_wait(tmem_full0 + 8u * acc, full_par);
      ptx::tcgen05_fence_after();
      if (et == 0) CSD_TSP(4);
      const uint32_t t_row = tmem_base + acc * kTPix + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * half_pix);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const int col0 = pass * 64, col1 = pass == 0 ? min(64, half_pix) : half_pix;
        if (col0 >= col1) break;                                  // uniform over the CTA
        if (lane == 0 && q == 0 && store_pending) ptx::bulk_wait_group_read0();   // (bulk groups are per thread)
        asm volatile("bar.sync 1, 256;" ::: "memory");            // staging tile free
#pragma unroll 1
        for (int col = col0; col < col1; col += 16) {
          uint32_t r0[16];
          __syncwarp();
          ptx::tmem_ld_x16(t_row + col, r0);
          // residual of the columns' pixels (column = 8 * tile row + x): issued before the TMEM wait
          float rv[16];
          if (res != nullptr && c_valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int pc = half * half_pix + col + i;
              const int gh = tc.h0 + (pc >> 3), gw = tc.w0 + (pc & 7);
              rv[i] = (col + i < m_lim_h) ? __ldg(res + (((long long)tc.b * p.H + gh) * p.W + gw) * p.res_pitch + c) : 0.f;
            }
          }
          ptx::tmem_ld_wait();
          if (c_valid) {
            float* sp = stage_h + (col - col0) * spitch;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float v = fmaf(__uint_as_float(r0[i]), scale, add_cs);
              if (res != nullptr) v = fmaf(rv[i], scale, v);
              if (col + i < m_lim_h) {
                s1 += v;
                s2 = fmaf(v, v, s2);
              }
              if (col + i < col1) sp[i * spitch] = v;
            }
          }
        }
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (lane == 0 && q == 0) {                                // one thread per half issues that half's store
          const int row0 = tc.h0 + half * half_rows + pass * 8;
          if (row0 < p.H) ptx::tma_store_4d(pass == 0 ? &mapOut : &mapOut2, stage_h_addr, tc.n0, tc.w0, row0, tc.b);
          ptx::bulk_commit_group();
        }
        if (et == 0) { if (pass == 0) CSD_TSP(6); else CSD_TSP(5); }
        store_pending = true;
      }
      ptx::tcgen05_fence_before();
      ptx::mbar_arrive(tmem_empty0 + 8u * acc);
      if (++acc == 2u) { acc = 0; full_par ^= 1u; }
      if (has_stats && c_valid) {
        float2* sp2 = reinterpret_cast<float2*>(p.stat_partials) + ((long long)(tc.sp * 2 + half) * p.n_store + c);
        *sp2 = make_float2(s1, s2);
      }
    }
    if (lane == 0 && q == 0 && store_pending) ptx::bulk_wait_group0();
  } else {
    // ===== epilogue (8 warps) =====
    // TMEM lane = output channel, column = pixel. Warp w reads lane quadrant w % 4 (the hardware's TMEM access
    // rule) and one pixel half. Each thread owns ONE channel: out = acc * scale + (bias + temb) * scale, its
    // GroupNorm partial sums (sum, sum of squares over the valid pixels) are plain register accumulations,
    // and the bf16 value goes to stage[pixel][channel] (a warp writes 32 consecutive channels of one pixel =
    // 64 contiguous bytes: conflict free). The accumulator is handed back to the MMA warp as soon as it has
    // been read; one elected thread then writes the staged 32 x 8 pixel tile to the NHWC output with a single
    // TMA store (rows beyond the image and channels beyond n_store are clipped by the TMA unit), which
    // drains while the warps wait for the next accumulator. Residual adds are not done here: the engine
    // appends the residual as an extra K segment with identity weights, so it rides on the tensor core.
    const int et = threadIdx.x - 320;                 // 0..255
    const int q = warp & 3;
    const int half = (warp - 10) >> 2;                // 0: pixels [0,128), 1: pixels [128,256)
    const int cl = q * 32 + lane;                     // channel inside the tile's 128-channel block
    const int spitch = p.out_box_c;                   // staging row pitch = channel extent of the TMA store box
    __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(__cvta_shared_to_generic(stage_base));
    const int half_pix = p.t_pix >> 1;                // pixels (accumulator columns) per epilogue half: 128 or 80
    __nv_bfloat16* stage_c = stage + (half * half_pix) * spitch + cl;
    const bool has_stats = p.stat_partials != nullptr;
    const float scale = p.scale;
    uint32_t acc = 0, full_par = 0;
    bool store_pending = false;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile);
      const int c = tc.n0 + cl;
      const int cb = min(kTChan, p.n_store - tc.n0);  // channels of this block that are stored (multiple of 8)
      const bool c_valid = cl < cb;
      const float add_cs = c_valid ? scale * ((p.bias != nullptr ? __ldg(p.bias + c) : 0.f) +
                                              (p.temb != nullptr ? __ldg(p.temb + (long long)tc.b * p.temb_pitch + c) : 0.f))
                                   : 0.f;
      // valid pixel columns of this half (rows beyond the image) and valid pixels per 8-pixel row (ragged width;
      // the engine only launches w % 8 == 0, where w_lim is always 8)
      const int w_lim = min(kHaloTW, p.W - tc.w0);
      const int m_lim_h = min(half_pix, min(p.t_pix, (p.H - tc.h0) * kHaloTW) - half * half_pix);
      const int m_lim = (w_lim == kHaloTW ? m_lim_h : 0);
      if (et == 0) CSD_TSP(3);
      ptx::mbar_wait(tmem_full0 + 8u * acc, full_par);
      ptx::tcgen05_fence_after();
      if (et == 0) CSD_TSP(4);
      if (et == 0 && store_pending) ptx::bulk_wait_group_read0();   // previous tile's store has left the staging tile
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const uint32_t t_row = tmem_base + acc * kTPix + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * half_pix);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int col = 0; col < ((p.debug_nodata & 32) ? 0 : half_pix); col += 32) {   // (probe bit 32: idle epilogue)
        uint32_t r0[16], r1[16];
        __syncwarp();
        ptx::tmem_ld_x16(t_row + col, r0);
        ptx::tmem_ld_x16(t_row + col + 16, r1);      // (a 80-column half reads 16 columns past its end: unused)
        ptx::tmem_ld_wait();
        if (c_valid) {
          __nv_bfloat16* sp = stage_c + col * spitch;
          if (col + 32 <= m_lim) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = fmaf(__uint_as_float(i < 16 ? r0[i] : r1[i - 16]), scale, add_cs);
              s1 += v;
              s2 = fmaf(v, v, s2);
              sp[i * spitch] = __float2bfloat16_rn(v);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = fmaf(__uint_as_float(i < 16 ? r0[i] : r1[i - 16]), scale, add_cs);
              if (col + i < m_lim_h && (i & 7) < w_lim) {
                s1 += v;
                s2 = fmaf(v, v, s2);
              }
              if (col + i < half_pix) sp[i * spitch] = __float2bfloat16_rn(v);   // never into the other half's rows
            }
          }
        }
      }
      ptx::tcgen05_fence_before();
      ptx::mbar_arrive(tmem_empty0 + 8u * acc);       // accumulator free: the next tile's MMAs may overwrite it
      if (++acc == 2u) { acc = 0; full_par ^= 1u; }
      if (has_stats && c_valid) {
        float2* sp2 = reinterpret_cast<float2*>(p.stat_partials) + ((long long)(tc.sp * 2 + half) * p.n_store + c);
        *sp2 = make_float2(s1, s2);
      }
      ptx::fence_proxy_async_smem();                  // staged tile -> visible to the TMA store (async proxy)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0) {
        CSD_TSP(5);
        ptx::tma_store_4d(&mapOut, stage_base, tc.n0, tc.w0, tc.h0, tc.b);
        ptx::bulk_commit_group();
      }
      store_pending = true;
    }
    if (et == 0 && store_pending) ptx::bulk_wait_group0();
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * kTPix);
  }
}

// Rows of the transposed kernel's macro tile (t_rows x 8 pixels = N of the MMA): the candidate that minimises
// tiles x cycles per instruction, with cycles ~ max(100, N / 2) (tools/mma_rate_probe.cu: an M = 128 instruction never
// costs less than ~100 cycles). 160 -> 32 (5 exact tiles), 80 -> 28 (3 tiles, 5 % padding instead of the 17 % of
// 3 x 32), 40 -> 20 (2 exact tiles). kernels.transposed_tile_rows() mirrors this rule.
static int pick_t_rows(int h) {
  if (const char* e = getenv("CSD_TROWS_LEGACY")) {      // A/B switch: the round-1 rule (32, or 20 for 40-row images)
    if (atoi(e) != 0) return (h % 32 != 0 && h < 64 && h % 20 == 0) ? 20 : 32;
  }
  const int cand[4] = {32, 28, 24, 20};
  int best = 32;
  long long best_cost = -1;
  for (int i = 0; i < 4; ++i) {
    const int t = cand[i];
    const long long cost = (long long)ceil_div(h, t) * std::max(100, t * 4);
    if (best_cost < 0 || cost <= best_cost) { best = t; best_cost = cost; }
  }
  return best;
}

static int next_pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

// Host launcher shared by csd_conv_gemm and the program executor (which pre-encodes the maps).
// ---------------------------------------------------------------------------------------------------
// Split-K finish (small levels). A 5 or 10 px level has fewer 128-pixel tiles than the GPU has SMs and every CTA
// walks the whole K = 9 * C range on its own, bound by the latency of its shared-memory ring rather than by the
// tensor core. With k_splits > 1 gridDim.z CTAs share one tile's K range, each writing raw fp32 partial sums to
// splitk_ws [split][pixel][ws_pitch]; this pass adds the partials in split order (fixed order: deterministic) and
// applies the epilogue the single-pass kernel would have: + bias[n] + temb[b][n] + res, * scale, optional tf32
// rounding, bf16 or fp32 store.
// ---------------------------------------------------------------------------------------------------
struct SplitKReduceParams {
  const float* ws;
  long long split_stride;   // floats between splits
  int ws_pitch, splits;
  long long pixels;
  int pix_per_image;
  int n_store;
  void* out; int out_pitch, out_f32, out_round;
  const float* bias;
  const float* temb; int temb_pitch;
  const void* res; int res_pitch, res_f32;
  float scale;
};

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const SplitKReduceParams p) {
  const int groups = (p.n_store + 3) >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.pixels * groups) return;
  const long long pix = idx / groups;
  const int n = (int)(idx - pix * groups) * 4;
  const float* w = p.ws + pix * p.ws_pitch + n;
  float4 acc = __ldg(reinterpret_cast<const float4*>(w));
  for (int s = 1; s < p.splits; ++s) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(w + s * p.split_stride));
    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
  }
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
  const int cnt = min(4, p.n_store - n);
  const int b = (int)(pix / p.pix_per_image);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i >= cnt) break;
    if (p.bias != nullptr) v[i] += __ldg(p.bias + n + i);
    if (p.temb != nullptr) v[i] += __ldg(p.temb + (long long)b * p.temb_pitch + n + i);
    if (p.res != nullptr)
      v[i] += p.res_f32 ? __ldg(reinterpret_cast<const float*>(p.res) + pix * p.res_pitch + n + i)
                        : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.res)[pix * p.res_pitch + n + i]);
    v[i] *= p.scale;
    if (p.out_round) v[i] = round_tf32(v[i]);
    if (p.out_f32) reinterpret_cast<float*>(p.out)[pix * p.out_pitch + n + i] = v[i];
    else reinterpret_cast<__nv_bfloat16*>(p.out)[pix * p.out_pitch + n + i] = __float2bfloat16_rn(v[i]);
  }
}

struct ConvGemmLaunch {
  int tap_chunk;   // channels per stage of the per-tap kernel (32 or 64)
  bool tf32;       // fp32 activations / weights, kind::tf32 (per-tap kernel only)
  CUtensorMap mapA[CSD_MAX_SEGMENTS];
  CUtensorMap mapB;
  CUtensorMap mapOut;
  CUtensorMap mapOut2;   // fp32 plan: store box of the remainder rows of a half tile (pass B)
  CUtensorMap mapW[CSD_MAX_SEGMENTS];   // transposed kernel: (k, row, tap) weight maps of the 9-tap segments
  ConvGemmKernelParams p;
  dim3 grid;
  size_t smem;
  bool halo;
  bool transposed;
  bool persistent;
  int ksplit;                // > 1: split-K partials + splitk_reduce_kernel
  SplitKReduceParams red;
};

int conv_gemm_prepare(const csd_conv_gemm_desc* d, ConvGemmLaunch* L) {
  CSD_REQUIRE(d != nullptr, "null conv_gemm desc");
  CSD_REQUIRE(d->nseg >= 1 && d->nseg <= CSD_MAX_SEGMENTS, "nseg=%d out of range", d->nseg);
  CSD_REQUIRE(d->batch >= 1 && d->h >= 1 && d->w >= 1, "bad spatial dims %d %d %d", d->batch, d->h, d->w);
  const bool t_mode = d->mode == 2;
  const bool halo_mode = d->mode == 1 || t_mode;
  const bool tf32 = d->dtype == 1;
  const int E = tf32 ? 4 : 2;                 // bytes per activation / weight element
  CSD_REQUIRE(d->dtype == 0 || d->dtype == 1, "conv_gemm: dtype=%d (0 = bf16, 1 = fp32 storage / tf32 operands)", d->dtype);
  CSD_REQUIRE(!tf32 || ((!halo_mode || t_mode) && d->out_f32 == 1),
              "conv_gemm: the fp32 / tf32 plan runs in the per-tap kernel (mode 0) or the transposed kernel (mode 2) and "
              "writes fp32");
  const int halo_chunk = tf32 ? 16 : kChunkK;     // channels per 64-byte halo chunk
  L->tf32 = tf32;
  L->halo = halo_mode;
  L->transposed = t_mode;
  L->persistent = false;
  const int mt = t_mode ? 2 : (halo_mode ? (d->mt > 0 ? d->mt : 1) : 1);
  CSD_REQUIRE(halo_mode || (d->tile_w >= 1 && d->tile_h >= 1 && d->tile_b >= 1 &&
                            d->tile_w * d->tile_h * d->tile_b <= kTileM),
              "tile box %dx%dx%d exceeds 128 pixels", d->tile_w, d->tile_h, d->tile_b);
  CSD_REQUIRE(!halo_mode || (d->z_batches == 1 && (d->stride <= 1) && d->pad == 1 && mt >= 1 && mt <= 4 &&
                             (t_mode || mt * d->n_tile <= 512)),
              "halo mode needs stride 1, pad 1, no z batching and mt*n_tile <= 512 (mt=%d n_tile=%d)", mt, d->n_tile);
  CSD_REQUIRE(!t_mode || (d->out_f32 == (tf32 ? 1 : 0) && d->bias_per_row == 0),
              "transposed halo mode writes bf16 (fp32 in the tf32 plan) with per-channel bias only");
  CSD_REQUIRE(d->n_tile >= 16 && d->n_tile % 16 == 0 && d->n_tile <= 512, "n_tile=%d invalid", d->n_tile);
  CSD_REQUIRE(d->n >= 1 && d->n_store >= 1, "n=%d n_store=%d invalid", d->n, d->n_store);
  CSD_REQUIRE(d->out != nullptr && d->wt != nullptr, "null out / wt pointer");
  CSD_REQUIRE(d->z_batches >= 1, "z_batches=%d", d->z_batches);

  ConvGemmKernelParams& p = L->p;
  memset(&p, 0, sizeof(p));
  p.B = d->batch; p.H = d->h; p.W = d->w;
  p.TW = halo_mode ? kHaloTW : d->tile_w;
  p.TH = halo_mode ? kHaloTH : d->tile_h;
  p.TB = halo_mode ? 1 : d->tile_b;
  p.mt = mt;
  // transposed mode: macro tile = t_rows x 8 pixels. 32 rows (N = 256) by default; 20 rows (N = 160) when that
  // tiles the image exactly and 32 would waste more than the shorter MMA loses (40-row images).
  // kernels.transposed_tile_rows() mirrors this rule for the statistics-partials geometry.
  p.t_rows = t_mode ? pick_t_rows(d->h) : kTRows;
  p.t_pix = p.t_rows * kHaloTW;
  p.tiles_w = ceil_div(d->w, p.TW);
  p.tiles_h = t_mode ? ceil_div(d->h, p.t_rows) : ceil_div(d->h, p.TH * mt);
  const int tiles_b = ceil_div(d->batch, p.TB);
  p.nseg = d->nseg;
  p.stride = d->stride > 0 ? d->stride : 1;
  p.pad = d->pad;
  CSD_REQUIRE(p.stride == 1 || p.stride == 2, "stride=%d unsupported", p.stride);
  CSD_REQUIRE(p.pad == 0 || p.pad == 1, "pad=%d unsupported", p.pad);
  const int in_h = d->in_h > 0 ? d->in_h : d->h, in_w = d->in_w > 0 ? d->in_w : d->w;
  p.n_store = d->n_store;
  p.n_tile = d->n_tile;
  if (d->n_tile <= 256) {
    p.nsplit = 1;
    p.n_sub = d->n_tile;
  } else {
    p.nsplit = 2;
    p.n_sub = d->n_tile / 2;
    CSD_REQUIRE(p.n_sub % 16 == 0, "n_tile=%d cannot be split into two multiples of 16", d->n_tile);
  }
  const int n_tiles = ceil_div(d->n_store, d->n_tile);
  CSD_REQUIRE(d->wt_rows >= 1, "wt_rows=%d", d->wt_rows);
  p.wt_k_off = d->wt_k_off;
  p.a_batch_step = d->a_batch_step;
  p.tmem_cols = t_mode ? kTPix : next_pow2_cols(d->n_tile * mt);

  // per-tap kernel: 64-channel stages unless every segment fits one 32-channel chunk (see conv_gemm_kernel)
  int tap_chunk = kChunkK;
  if (!halo_mode && !tf32 && getenv("CSD_TAP_CHUNK32") == nullptr)
    for (int s = 0; s < d->nseg; ++s)
      if (d->seg[s].c_cnt > kChunkK) tap_chunk = 64;
  L->tap_chunk = tap_chunk;
  const int tap_row_bytes = tap_chunk * E;     // 128 (SWIZZLE_128B: 64 bf16 or 32 fp32 channels) or 64
  const CUtensorMapDataType tm_dtype = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  int tap_iters = 0;
  int k_total = 0;
  int k_total_chan = 0;  // padded channels over all segments (one (scale, shift) slot each)
  for (int s = 0; s < d->nseg; ++s) {
    const csd_conv_segment& sg = d->seg[s];
    k_total_chan += ceil_div(sg.c_cnt, halo_chunk) * halo_chunk;
    CSD_REQUIRE(sg.a != nullptr, "segment %d: null tensor", s);
    CSD_REQUIRE(sg.taps == 1 || sg.taps == 9, "segment %d: taps=%d (1 or 9)", s, sg.taps);
    CSD_REQUIRE(sg.pitch % 8 == 0 && sg.c_cnt >= 1 && sg.c_off >= 0 && sg.c_off + sg.c_cnt <= sg.pitch,
                "segment %d: bad channel range off=%d cnt=%d pitch=%d", s, sg.c_off, sg.c_cnt, sg.pitch);
    p.seg_taps[s] = sg.taps;
    p.seg_chunks[s] = ceil_div(sg.c_cnt, halo_mode ? halo_chunk : tap_chunk);
    p.seg_coff[s] = sg.c_off;
    p.seg_kbase[s] = k_total / kChunkK;
    k_total += sg.taps * ceil_div(sg.c_cnt, kChunkK) * kChunkK;
    tap_iters += sg.taps * p.seg_chunks[s];
    CSD_REQUIRE(sg.norm == nullptr || (t_mode && sg.c_off == 0),
                "segment %d: the fused GroupNorm prologue needs the transposed halo mode and c_off == 0", s);
    p.seg_norm[s] = sg.norm;
    p.seg_silu[s] = sg.norm_silu;
    p.seg_ccnt[s] = sg.c_cnt;
    if (sg.norm != nullptr) p.has_norm = 1;
    // 4-D map over [batch, h, w, c]; dim 0 stops at the last valid channel so the remainder of a
    // 32-channel chunk is zero-filled instead of reading the neighbouring channels.
    const uint64_t z_extra = (uint64_t)(d->z_batches - 1) * (uint64_t)d->a_batch_step;
    uint64_t dims[4] = {(uint64_t)(sg.c_off + sg.c_cnt), (uint64_t)in_w, (uint64_t)in_h,
                        (uint64_t)d->batch + z_extra};
    uint64_t strides[3] = {(uint64_t)sg.pitch * E, (uint64_t)sg.pitch * E * in_w,
                           (uint64_t)sg.pitch * E * in_w * in_h};
    // with a traversal stride s the box spans (t-1)*s+1 source elements and delivers t of them
    uint32_t box[4] = {(uint32_t)(halo_mode ? halo_chunk : tap_chunk), (uint32_t)((p.TW - 1) * p.stride + 1),
                       (uint32_t)((p.TH - 1) * p.stride + 1), (uint32_t)p.TB};
    if (halo_mode) {  // whole halo of the mt stacked tiles (a 1-tap segment needs no halo)
      const int hl = sg.taps == 9 ? 1 : 0;
      box[1] = (uint32_t)(kHaloTW + 2 * hl);
      box[2] = (uint32_t)((t_mode ? p.t_rows : kHaloTH * mt) + 2 * hl);
    }
    uint32_t estr[4] = {1, (uint32_t)p.stride, (uint32_t)p.stride, 1};
    int st = encode_tensor_map(&L->mapA[s], tm_dtype, 4, sg.a, dims, strides, box,
                               (!halo_mode && tap_row_bytes == 128) ? TMA_SW_128 : TMA_SW_64, estr);
    if (st != CSD_OK) return st;
  }
  for (int s = d->nseg; s < CSD_MAX_SEGMENTS; ++s) L->mapA[s] = L->mapA[0];
  CSD_REQUIRE(k_total == d->k_total, "k_total mismatch: segments give %d, desc says %d", k_total, d->k_total);
  {
    // 3-D map over Wt [z][rows][k]; rows beyond wt_rows and k beyond k_total are zero-filled.
    const uint64_t zb = d->wt_batch_stride != 0 ? (uint64_t)d->z_batches : 1;
    uint64_t dims[3] = {(uint64_t)(d->wt_k_off + (d->k_valid > 0 ? d->k_valid : d->k_total)), (uint64_t)d->wt_rows, zb};
    const uint64_t row_bytes = (uint64_t)(d->wt_pitch > 0 ? d->wt_pitch : d->k_total) * E;
    uint64_t strides[2] = {row_bytes, d->wt_batch_stride != 0 ? (uint64_t)d->wt_batch_stride * E
                                                               : row_bytes * (uint64_t)d->wt_rows};
    // (a slab box that stops at the last stored channel - 96 instead of 128 rows, a quarter less weight fill - was
    //  measured neutral: 28.79 vs 28.68 ms per step; the full box stays)
    p.b_box_rows = kTChan;
    if (const char* e = getenv("CSD_DEBUG_SLAB_ROWS")) p.b_box_rows = std::max(8, std::min(kTChan, atoi(e)));   // probe: wrong results
    uint32_t box[3] = {(uint32_t)(halo_mode ? halo_chunk : tap_chunk), (uint32_t)(t_mode ? p.b_box_rows : p.n_sub), 1};
    int st = encode_tensor_map(&L->mapB, tm_dtype, 3, d->wt, dims, strides, box,
                               (!halo_mode && tap_row_bytes == 128) ? TMA_SW_128 : TMA_SW_64);
    if (st != CSD_OK) return st;
  }

  p.a_box_bytes = (uint32_t)(p.TW * p.TH * p.TB * tap_row_bytes);
  p.b_box_bytes = (uint32_t)(p.n_sub * tap_row_bytes);
  p.stage_bytes = (uint32_t)((kTileM * tap_row_bytes + d->n_tile * tap_row_bytes + 1023) & ~1023);
  const int ks = d->k_splits > 1 ? d->k_splits : 1;
  CSD_REQUIRE(ks == 1 || (!halo_mode && d->z_batches == 1 && d->bias_per_row == 0 && d->splitk_ws != nullptr &&
                          ks <= 16 && ks <= tap_iters),
              "k_splits=%d needs the per-tap kernel (mode 0), z_batches == 1, per-channel bias, a workspace and at "
              "least one K iteration per split (%d iterations)", ks, tap_iters);
  L->ksplit = ks;
  const int total_iters = ceil_div(tap_iters, ks);
  int budget = p.tmem_cols <= 128 ? 56 * 1024 : (p.tmem_cols <= 256 ? 100 * 1024 : 200 * 1024);
  {
    // Small grids (the 5/10/20 px levels: fewer CTAs than SMs) run one CTA per SM whatever their footprint and
    // their K loop is TMA-latency bound, so they take the whole shared memory for a deeper ring instead of
    // leaving room for co-resident CTAs that will never come.
    const long long ctas = (long long)p.tiles_w * p.tiles_h * tiles_b * n_tiles * d->z_batches * ks;
    if (!halo_mode && ctas <= num_sms()) budget = 200 * 1024;
  }
  int stages = budget / (int)p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (const char* e = getenv("CSD_TAP_STAGES")) stages = std::min(stages, std::max(2, atoi(e)));   // probe only
  if (stages > total_iters) stages = total_iters;
  if (stages < 2) stages = 2;
  if (stages >= kTapProducers) {
    stages -= stages % kTapProducers;
    p.producers = kTapProducers;
  } else {
    p.producers = stages;
  }
  p.num_stages = stages;

  p.out = d->out; p.out_pitch = d->out_pitch; p.out_f32 = d->out_f32; p.out_z_stride = d->out_z_stride;
  p.out_round = tf32 ? d->out_round_tf32 : 0;
  p.bias = d->bias; p.bias_per_row = d->bias_per_row;
  p.temb = d->temb; p.temb_pitch = d->temb_pitch;
  p.res = reinterpret_cast<const __nv_bfloat16*>(d->res); p.res_pitch = d->res_pitch;
  p.res_z_stride = d->res_z_stride;
  p.scale = d->scale;
  p.ksplit = ks;
  if (ks > 1) {
    // the conv kernel writes raw partial sums; the epilogue moves to splitk_reduce_kernel
    SplitKReduceParams& r = L->red;
    const int ws_pitch = ceil_div(d->n_store, 8) * 8;
    r.ws = d->splitk_ws; r.ws_pitch = ws_pitch; r.splits = ks;
    r.pixels = (long long)d->batch * d->h * d->w; r.pix_per_image = d->h * d->w;
    r.split_stride = r.pixels * ws_pitch;
    r.n_store = d->n_store;
    r.out = d->out; r.out_pitch = d->out_pitch; r.out_f32 = (tf32 || d->out_f32) ? 1 : 0; r.out_round = p.out_round;
    r.bias = d->bias; r.temb = d->temb; r.temb_pitch = d->temb_pitch;
    r.res = d->res; r.res_pitch = d->res_pitch; r.res_f32 = tf32 ? 1 : 0;
    r.scale = d->scale;
    p.out = d->splitk_ws; p.out_pitch = ws_pitch; p.out_f32 = 1; p.out_z_stride = r.split_stride;
    p.out_round = 0; p.bias = nullptr; p.temb = nullptr; p.res = nullptr; p.scale = 1.0f;
  }
  {
    const char* e = getenv("CSD_DEBUG_NODATA");
    p.debug_nodata = (e != nullptr) ? atoi(e) : 0;  // bit 0: skip weight loads, bit 1: skip activation loads
    const char* t = getenv("CSD_DEBUG_TS");
    p.debug_ts = (t != nullptr) ? reinterpret_cast<long long*>(strtoull(t, nullptr, 0)) : nullptr;
  }
  CSD_REQUIRE(d->out_pitch % 8 == 0, "out_pitch=%d must be a multiple of 8", d->out_pitch);
  CSD_REQUIRE(d->res == nullptr || d->res_pitch % 8 == 0, "res_pitch=%d must be a multiple of 8", d->res_pitch);

  L->grid = dim3((unsigned)(p.tiles_w * p.tiles_h * tiles_b), (unsigned)n_tiles, (unsigned)(d->z_batches * ks));
  L->smem = (size_t)stages * p.stage_bytes + 1024 /*alignment slack*/ + 8 * (2 * kMaxStages + 2) + 4 * kAddendFloats;
  if (halo_mode) {
    p.a_stage_bytes = (uint32_t)(((kHaloTW + 2) * (kHaloTH * mt + 2) * kRowBytes + 1023) & ~1023);
    p.b_stage_bytes = (uint32_t)(((t_mode ? kTChan : d->n_tile) * kRowBytes + 1023) & ~1023);
    // two CTAs per SM when TMEM allows it (<= 256 columns): keep each under ~110 KB
    const int budget_h = p.tmem_cols <= 256 ? 108 * 1024 : 200 * 1024;
    p.a_stages = 2;
    int bs = (budget_h - p.a_stages * (int)p.a_stage_bytes) / (int)p.b_stage_bytes;
    if (bs > kMaxBStages) bs = kMaxBStages;
    // A halo lasts 9 taps, so two halo buffers already cover its load latency; the weight ring is the
    // latency-critical one (one slab per tap) and gets the rest of the budget. A third halo buffer is
    // only taken when the weight ring is already at its maximum depth.
    if (bs == kMaxBStages && budget_h - 3 * (int)p.a_stage_bytes >= kMaxBStages * (int)p.b_stage_bytes) p.a_stages = 3;
    CSD_REQUIRE(bs >= 2, "halo mode: not enough shared memory for the weight ring (n_tile=%d mt=%d)", d->n_tile, mt);
    size_t tail = 8 * (2 * kMaxAStages + 2 * kMaxBStages + 2) + 4 * 512;
    if (t_mode && p.has_norm) {
      // TMA -> transform -> MMA: three halo buffers (one loading, one being normalised, one feeding the tensor
      // core), the a_ready barriers and the (scale, shift) table of every K channel
      tail = 8 * (2 * kMaxAStages + 2 * kMaxBStages + 2) + 8 * kMaxAStages + (size_t)k_total_chan * 8 + 16;
      p.a_stages = 3;
      bs = ((int)(113 * 1024 - 1024 - tail) - p.a_stages * (int)p.a_stage_bytes) / (int)p.b_stage_bytes;
      if (bs > kMaxBStages) bs = kMaxBStages;
      CSD_REQUIRE(bs >= 3, "fused GroupNorm prologue: not enough shared memory for the weight ring");
    }
    p.b_stages = bs;
    L->smem = (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * p.b_stage_bytes + 1024 + tail;
    if (t_mode) L->grid.y = (unsigned)ceil_div(d->n_store, kTChan);
    L->persistent = false;
    if (t_mode) {
      // chunk schedule: 1-tap chunks spread evenly behind the 9-tap chunks (see ConvGemmKernelParams::sched_seg)
      {
        std::vector<std::pair<int, int>> nine, one, order;
        int tab = 0;
        for (int s = 0; s < d->nseg; ++s) {
          p.seg_tab_base[s] = tab;
          tab += p.seg_chunks[s] * halo_chunk;
          for (int c = 0; c < p.seg_chunks[s]; ++c) (p.seg_taps[s] == 9 ? nine : one).push_back({s, c});
        }
        p.n_sched = (int)(nine.size() + one.size());
        const bool interleave = !nine.empty() && !one.empty() && p.n_sched <= kMaxSched &&
                                getenv("CSD_NO_CHUNK_INTERLEAVE") == nullptr;
        if (interleave) {
          const size_t n9 = nine.size(), n1 = one.size();
          for (size_t i = 0; i < n9; ++i) {      // (1-tap chunks in front of their 9-tap chunk measured the same)
            order.push_back(nine[i]);
            for (size_t j = i * n1 / n9; j < (i + 1) * n1 / n9; ++j) order.push_back(one[j]);
          }
          for (int i = 0; i < p.n_sched; ++i) {
            p.sched_seg[i] = (unsigned char)order[i].first;
            p.sched_chunk[i] = (unsigned char)order[i].second;
          }
        }
        p.sched_tab = interleave ? 1 : 0;
      }
      // persistent kernel: one CTA per SM, whole shared memory
      L->persistent = true;
      p.n_blocks = ceil_div(d->n_store, kTChan);
      p.num_tiles = p.tiles_w * p.tiles_h * tiles_b * p.n_blocks;
      // with the fused prologue a halo goes TMA -> transform -> MMA: a fourth buffer lets the next tile's first
      // chunk be fetched and normalised while the current tile still has two chunks to multiply
      p.a_stages = p.has_norm ? 4 : 3;
      if (const char* e = getenv("CSD_TP_A_STAGES")) p.a_stages = std::max(2, std::min(kMaxAStages, atoi(e)));   // probe
      // (Halo buffers and staging tile sized per launch - 15 instead of 9 weight slabs in flight at 40 px - measured
      //  neutral, like quartering the slab bytes and doubling the producer warps: the ~15-20 % the weight stream costs at
      //  N = 160 / 224 is neither bytes, nor ring depth, nor TMA issue rate; profiles/conv_nodata_r2.txt.)
      p.staging_bytes = kPStagingBytes;
      if (getenv("CSD_FIXED_TP_SMEM") == nullptr) {
        // halo buffers and staging tile sized for this launch's macro tile: at 40 px (20-row tiles) that is room for
        // five 3-slab ring slots instead of three
        p.out_box_c = std::min(kTChan, d->n_store);
        p.a_stage_bytes = (uint32_t)(((kHaloTW + 2) * (p.t_rows + 2) * kRowBytes + 511) & ~511);   // SW64 pattern period
        const int stg = tf32 ? 512 * p.out_box_c : 2 * p.t_pix * p.out_box_c;
        p.staging_bytes = (stg + 1023) & ~1023;
      }
      // weight ring: slots of three slabs (the three taps of a kernel row, one TMA op through mapW[s])
      p.b_stage_bytes = 3u * kTChan * kRowBytes;
      const size_t fixed = 1024 + (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.staging_bytes + kPBarBytes +
                           (size_t)k_total_chan * 8 + 16;
      CSD_REQUIRE(fixed + 2 * (size_t)p.b_stage_bytes <= 227 * 1024, "transposed conv: K=%d channels too many for the "
                  "shared-memory coefficient table", k_total_chan);
      int pbs = (int)((227 * 1024 - fixed) / p.b_stage_bytes);
      if (pbs > kPMaxBStages) pbs = kPMaxBStages;
      p.b_stages = pbs;
      for (int s = 0; s < CSD_MAX_SEGMENTS; ++s) L->mapW[s] = L->mapB;
      for (int s = 0; s < d->nseg; ++s) {
        if (p.seg_taps[s] != 9) continue;
        // (k, row, tap) view of this segment's K range: Wt[row][wt_k_off + seg_kbase * 32 + tap * kstep + k]
        const int kstep = ceil_div(p.seg_ccnt[s], kChunkK) * kChunkK;
        const uint64_t row_bytes = (uint64_t)(d->wt_pitch > 0 ? d->wt_pitch : d->k_total) * E;
        const char* base = static_cast<const char*>(d->wt) + ((size_t)d->wt_k_off + (size_t)p.seg_kbase[s] * kChunkK) * E;
        uint64_t wdims[3] = {(uint64_t)kstep, (uint64_t)d->wt_rows, 9};
        uint64_t wstr[2] = {row_bytes, (uint64_t)kstep * E};
        uint32_t wbox[3] = {(uint32_t)halo_chunk, (uint32_t)kTChan, 3};
        int st = encode_tensor_map(&L->mapW[s], tm_dtype, 3, base, wdims, wstr, wbox, TMA_SW_64);
        if (st != CSD_OK) return st;
      }
      L->smem = fixed + (size_t)pbs * p.b_stage_bytes;
      const int sms = num_sms();
      L->grid = dim3((unsigned)std::min(p.num_tiles, sms), 1, 1);
      CSD_REQUIRE(d->res == nullptr || tf32,
                  "transposed conv (bf16): pass the residual as a 1-tap segment with identity weights");
      CSD_REQUIRE(d->res == nullptr || d->res_pitch >= d->n_store, "transposed conv: residual pitch %d", d->res_pitch);
      // TMA store map over out [batch, h, w, out_pitch], box = 128 (or n_store) channels x 8 x 32 pixels; the fp32
      // plan stores 32 channels per pass
      p.out_box_c = std::min(kTChan, d->n_store);
      uint64_t odims[4] = {(uint64_t)d->n_store, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->batch};
      uint64_t ostr[3] = {(uint64_t)d->out_pitch * E, (uint64_t)d->out_pitch * E * d->w,
                          (uint64_t)d->out_pitch * E * d->w * d->h};
      uint32_t obox[4] = {(uint32_t)p.out_box_c, (uint32_t)kHaloTW, (uint32_t)(tf32 ? 8 : p.t_rows), 1};
      int st = encode_tensor_map(&L->mapOut, tm_dtype, 4, d->out, odims, ostr, obox, TMA_SW_NONE);
      if (st != CSD_OK) return st;
      L->mapOut2 = L->mapOut;
      if (tf32 && p.t_rows / 2 > 8) {      // pass B of the fp32 epilogue: the half tile's rows beyond the first 8
        uint32_t obox2[4] = {(uint32_t)p.out_box_c, (uint32_t)kHaloTW, (uint32_t)(p.t_rows / 2 - 8), 1};
        st = encode_tensor_map(&L->mapOut2, tm_dtype, 4, d->out, odims, ostr, obox2, TMA_SW_NONE);
        if (st != CSD_OK) return st;
      }
    }
  }
  p.stat_partials = t_mode ? d->stat_partials : nullptr;
  CSD_REQUIRE(d->stat_partials == nullptr || t_mode, "stat_partials are produced by the transposed halo mode only");
  CSD_REQUIRE(L->smem <= 227 * 1024, "shared memory %zu exceeds 227 KB", L->smem);
  // If z batches share the weights the z coordinate of the weight map must stay 0.
  if (d->wt_batch_stride == 0 && d->z_batches > 1) {
    return set_error(CSD_ERR_UNSUPPORTED, "z_batches > 1 requires wt_batch_stride != 0");
  }
  return CSD_OK;
}

int conv_gemm_launch(const ConvGemmLaunch* L, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CSD_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CSD_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CSD_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CSD_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CSD_CUDA(cudaFuncSetAttribute(conv_halo_tp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CSD_CUDA(cudaFuncSetAttribute(conv_halo_tp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (L->persistent && L->tf32) {
    conv_halo_tp_kernel<true><<<L->grid, kPThreads, L->smem, stream>>>(L->mapA[0], L->mapA[1], L->mapA[2], L->mapA[3],
                                                                         L->mapB, L->mapOut, L->mapOut2, L->mapW[0], L->mapW[1], L->mapW[2],
                                                                         L->mapW[3], L->p);
  } else if (L->persistent) {
    conv_halo_tp_kernel<false><<<L->grid, kPThreads, L->smem, stream>>>(L->mapA[0], L->mapA[1], L->mapA[2], L->mapA[3],
                                                                          L->mapB, L->mapOut, L->mapOut, L->mapW[0], L->mapW[1], L->mapW[2],
                                                                          L->mapW[3], L->p);
  } else if (L->halo) {
    conv_halo_kernel<<<L->grid, kConvThreads, L->smem, stream>>>(L->mapA[0], L->mapA[1], L->mapA[2], L->mapA[3],
                                                                 L->mapB, L->p);
  } else if (L->tf32) {
    conv_gemm_kernel<32, true><<<L->grid, kTapThreads, L->smem, stream>>>(L->mapA[0], L->mapA[1], L->mapA[2], L->mapA[3],
                                                                          L->mapB, L->p);
  } else if (L->tap_chunk == 64) {
    conv_gemm_kernel<64, false><<<L->grid, kTapThreads, L->smem, stream>>>(L->mapA[0], L->mapA[1], L->mapA[2],
                                                                           L->mapA[3], L->mapB, L->p);
  } else {
    conv_gemm_kernel<32, false><<<L->grid, kTapThreads, L->smem, stream>>>(L->mapA[0], L->mapA[1], L->mapA[2],
                                                                           L->mapA[3], L->mapB, L->p);
  }
  CSD_LAUNCH_CHECK("conv_gemm_kernel");
  if (L->ksplit > 1) {
    const long long work = L->red.pixels * ((L->red.n_store + 3) / 4);
    splitk_reduce_kernel<<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(L->red);
    CSD_LAUNCH_CHECK("splitk_reduce_kernel");
  }
  return CSD_OK;
}

}  // namespace csd

extern "C" int csd_conv_gemm(const csd_conv_gemm_desc* desc, csd_stream_t stream) {
  csd::ConvGemmLaunch L;
  int st = csd::conv_gemm_prepare(desc, &L);
  if (st != CSD_OK) return st;
  return csd::conv_gemm_launch(&L, static_cast<cudaStream_t>(stream));
}
