// Fused self-attention core on tcgen05 (sm_100a): S = Q K^T -> softmax -> O = P V -> Y = O Wo + bo -> (x + Y) * scale,
// ONE kernel per attention block, one CTA per (image, 128-query tile). The logits S, the probabilities P and the
// attention output O never exist in HBM (the reference - and this library's round-1 path - materialise [B, L, L] twice).
//
// Reference: AttnBlockpp.forward (models/layerspp.py:75-91) and the DDPM AttnBlock (models/layers.py:577-590):
//   h = GroupNorm(x); q, k, v = NIN_0/1/2(h); w = softmax(einsum('bchw,bcij->bhwij', q, k) * C^-0.5 over ij);
//   h = einsum('bhwij,bcij->bchw', w, v); h = NIN_3(h); return (x + h) / sqrt(2)   [plain x + h for the DDPM block]
// GroupNorm and the q|k|v projection stay separate launches (gn_fused + one 1x1 csd_conv_gemm writing [B, L, 3C]); this
// kernel replaces the two batched GEMMs, the softmax pass and the output projection.
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2-9 = softmax /
// conversion / epilogue (thread <-> query row = TMEM lane; the two warps that share a lane quadrant split the columns).
// Phases of one CTA, all operands bf16 in 128-byte-swizzled shared memory, accumulators fp32 in TMEM:
//   1. S[128 q, L keys] = sum over 64-channel chunks of Q_c K_c^T          (A, B K-major; N = keys in chunks of <= 256)
//   2. softmax warps: row max, p = exp2((s - max) * scale * log2 e) rounded to bf16, row sums of the ROUNDED values;
//      P is written to shared memory in the K-major operand layout (normalisation by 1/sum is deferred to step 4)
//   3. O[128 q, C] = sum over 64-key chunks of P_j V_j                     (B = V MN-major straight from [B, L, 3C])
//   4. softmax warps: O * (1 / row sum) -> bf16 -> shared memory (K-major A operand of the projection)
//   5. Y[128 q, C] = sum over 64-channel chunks of O_c Wo_c^T              (Wo = packed NIN_3 weights, K-major)
//   6. epilogue: out = (Y + bo + x) * scale, bf16 NHWC rows
// TMEM: S, O and Y reuse columns [0, 512): S is dead once P is in shared memory, O once its bf16 copy is.
// Every mbarrier is used for exactly one CTA tile (the kernel is not persistent), so ring parities are plain counters.
#include "common.cuh"
#include "ptx.cuh"
#include "tensormap.cuh"
#include "../../include/csd_b200.h"

#include <algorithm>

namespace csd {

constexpr int kAtThreads = 320;
constexpr int kAtEpi = 256;             // softmax / epilogue threads (warps 2..9)
constexpr int kAtQ = 128;               // query rows per CTA (UMMA M)
constexpr uint32_t kAtSw128 = 2;        // UMMA layout type SWIZZLE_128B
constexpr int kAtMaxRing = 4;

struct AttnParams {
  int L, C, batch;
  int nc;                // 64-channel chunks of C
  int nk;                // 64-key chunks of L
  int kbox;              // 128-row K boxes per chunk = ceil(L / 128)
  int lp;                // L rounded up to 16: S columns / issued key steps
  uint32_t a_slot, off_p, off_vw, vw_slot, bar_off;   // shared-memory map (bytes from the 1 KB aligned base)
  int a_slots, vw_slots;
  float scale_log2e;     // C^-0.5 * log2(e)
  float out_scale;       // 1/sqrt(2) (AttnBlockpp with skip_rescale) or 1
  const float* bo;
  const __nv_bfloat16* res;
  int res_pitch;
  __nv_bfloat16* out;
  int out_pitch;
};

// element (row, col) of a K-major bf16 operand stored as 64-column blocks of [128 rows x 128 bytes] with the hardware's
// 128-byte swizzle (16-byte unit index XOR row % 8; blocks are 1 KB aligned, so address bits = offset bits)
__device__ __forceinline__ uint32_t kmajor_unit_off(int row, int col /* multiple of 8 */) {
  const int blk = col >> 6, u = (col & 63) >> 3;
  return (uint32_t)(blk * (kAtQ * 128) + row * 128 + ((u ^ (row & 7)) << 4));
}

__global__ void __launch_bounds__(kAtThreads, 1)
attn_core_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapW, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar = smem_base + p.bar_off;
  // barriers: a_full[4] a_empty[4] vw_full[4] vw_empty[4] s_full p_ready o_full o_ready y_full | tmem slot | exchange
  const uint32_t a_full = bar, a_empty = bar + 32, vw_full = bar + 64, vw_empty = bar + 96;
  const uint32_t s_full = bar + 128, p_ready = bar + 136, o_full = bar + 144, o_ready = bar + 152, y_full = bar + 160;
  const uint32_t tmem_slot = bar + 168;
  float* xch = reinterpret_cast<float*>(smem_gen + p.bar_off + 176);   // [2 halves][128 rows] row max, then row sum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAtQ, b = blockIdx.y;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapQ);
    ptx::prefetch_tensormap(&mapK);
    ptx::prefetch_tensormap(&mapV);
    ptx::prefetch_tensormap(&mapW);
    for (int s = 0; s < kAtMaxRing; ++s) {
      ptx::mbar_init(a_full + 8u * s, 1);
      ptx::mbar_init(a_empty + 8u * s, 1);
      ptx::mbar_init(vw_full + 8u * s, 1);
      ptx::mbar_init(vw_empty + 8u * s, 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_ready, kAtEpi);
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_ready, kAtEpi);
    ptx::mbar_init(y_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int nblk = p.nc;                          // 64-channel blocks of V / Wo per ring item
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      // phase 1 ring: (Q chunk, K chunk)
      const uint32_t a_bytes = (uint32_t)(kAtQ * 128 + p.kbox * 128 * 128);
      for (int c = 0; c < p.nc; ++c) {
        const int s = c % p.a_slots;
        ptx::mbar_wait(a_empty + 8u * s, ((c / p.a_slots) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(a_full + 8u * s, a_bytes);
        const uint32_t dst = smem_base + s * p.a_slot;
        ptx::tma_load_3d(dst, &mapQ, a_full + 8u * s, c * 64, q0, b);
        for (int kb = 0; kb < p.kbox; ++kb)
          ptx::tma_load_3d(dst + kAtQ * 128 + kb * (128 * 128), &mapK, a_full + 8u * s, c * 64, kb * 128, b);
      }
      // phases 3 and 5 share one ring: V key chunks, then Wo channel chunks (each item = nblk boxes of 64 x 64)
      const uint32_t vw_bytes = (uint32_t)(nblk * 64 * 128);
      const int items = p.nk + p.nc;
      for (int it = 0; it < items; ++it) {
        const int s = it % p.vw_slots;
        ptx::mbar_wait(vw_empty + 8u * s, ((it / p.vw_slots) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(vw_full + 8u * s, vw_bytes);
        const uint32_t dst = smem_base + p.off_vw + s * p.vw_slot;
        if (it < p.nk) {
          for (int blk = 0; blk < nblk; ++blk)       // V[keys 64*it.., channels 64*blk..]: 64 key rows x 128 bytes
            ptx::tma_load_3d(dst + blk * (64 * 128), &mapV, vw_full + 8u * s, blk * 64, it * 64, b);
        } else {
          const int c = it - p.nk;
          for (int blk = 0; blk < nblk; ++blk)       // Wo[out rows 64*blk.., k columns 64*c..]: 64 rows x 128 bytes
            ptx::tma_load_3d(dst + blk * (64 * 128), &mapW, vw_full + 8u * s, c * 64, blk * 64, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t hi = ptx::smem_desc_hi(1024, kAtSw128);
      // ---- phase 1: S = Q K^T ----
      const int n0 = min(p.lp, 256), n1 = p.lp - n0;              // key columns of the two accumulator halves
      const uint32_t id0 = ptx::make_idesc_bf16_m128((uint32_t)n0);
      const uint32_t id1 = n1 > 0 ? ptx::make_idesc_bf16_m128((uint32_t)n1) : 0u;
      uint32_t acc = 0;
      for (int c = 0; c < p.nc; ++c) {
        const int s = c % p.a_slots;
        ptx::mbar_wait(a_full + 8u * s, (c / p.a_slots) & 1);
        ptx::tcgen05_fence_after();
        const uint32_t q_lo = ptx::smem_desc_lo(smem_base + s * p.a_slot, 16);
        const uint32_t k_lo = ptx::smem_desc_lo(smem_base + s * p.a_slot + kAtQ * 128, 16);
        const int ksteps = min(4, (p.C - c * 64 + 15) >> 4);
        for (int k16 = 0; k16 < ksteps; ++k16) {
          ptx::mma_bf16_ss(tmem_base, ptx::smem_desc_join(hi, q_lo + 2 * k16), ptx::smem_desc_join(hi, k_lo + 2 * k16), id0,
                           acc);
          if (n1 > 0)
            ptx::mma_bf16_ss(tmem_base + 256, ptx::smem_desc_join(hi, q_lo + 2 * k16),
                             ptx::smem_desc_join(hi, k_lo + ((256 * 128) >> 4) + 2 * k16), id1, acc);
          acc = 1u;
        }
        ptx::mma_commit(a_empty + 8u * s);
      }
      ptx::mma_commit(s_full);
      // ---- phase 3: O = P V (B = V MN-major: 64-channel blocks 8 KB apart, 16 key rows = 2 KB per K step) ----
      ptx::mbar_wait(p_ready, 0);
      ptx::tcgen05_fence_after();
      const int cn0 = p.C <= 256 ? p.C : 192, cn1 = p.C - cn0;    // output channels per MMA (N <= 256)
      const uint32_t idv0 = ptx::make_idesc_bf16_m128((uint32_t)cn0) | (1u << 16);
      const uint32_t idv1 = cn1 > 0 ? (ptx::make_idesc_bf16_m128((uint32_t)cn1) | (1u << 16)) : 0u;
      acc = 0;
      int it = 0;
      for (int j = 0; j < p.nk; ++j, ++it) {
        const int s = it % p.vw_slots;
        ptx::mbar_wait(vw_full + 8u * s, (it / p.vw_slots) & 1);
        ptx::tcgen05_fence_after();
        const uint32_t p_lo = ptx::smem_desc_lo(smem_base + p.off_p + j * (kAtQ * 128), 16);
        const uint32_t v_lo = ptx::smem_desc_lo(smem_base + p.off_vw + s * p.vw_slot, 64 * 128);
        const int ksteps = min(4, (p.lp - j * 64) >> 4);
        for (int k16 = 0; k16 < ksteps; ++k16) {
          ptx::mma_bf16_ss(tmem_base, ptx::smem_desc_join(hi, p_lo + 2 * k16), ptx::smem_desc_join(hi, v_lo + 128 * k16), idv0,
                           acc);
          if (cn1 > 0)
            ptx::mma_bf16_ss(tmem_base + 192, ptx::smem_desc_join(hi, p_lo + 2 * k16),
                             ptx::smem_desc_join(hi, v_lo + ((3 * 64 * 128) >> 4) + 128 * k16), idv1, acc);
          acc = 1u;
        }
        ptx::mma_commit(vw_empty + 8u * s);
      }
      ptx::mma_commit(o_full);
      // ---- phase 5: Y = O Wo^T (both K-major; Wo rows = output channels) ----
      ptx::mbar_wait(o_ready, 0);
      ptx::tcgen05_fence_after();
      const uint32_t idw0 = ptx::make_idesc_bf16_m128((uint32_t)cn0);
      const uint32_t idw1 = cn1 > 0 ? ptx::make_idesc_bf16_m128((uint32_t)cn1) : 0u;
      acc = 0;
      for (int c = 0; c < p.nc; ++c, ++it) {
        const int s = it % p.vw_slots;
        ptx::mbar_wait(vw_full + 8u * s, (it / p.vw_slots) & 1);
        ptx::tcgen05_fence_after();
        const uint32_t o_lo = ptx::smem_desc_lo(smem_base + p.off_p + c * (kAtQ * 128), 16);
        const uint32_t w_lo = ptx::smem_desc_lo(smem_base + p.off_vw + s * p.vw_slot, 16);
        const int ksteps = min(4, (p.C - c * 64 + 15) >> 4);
        for (int k16 = 0; k16 < ksteps; ++k16) {
          ptx::mma_bf16_ss(tmem_base, ptx::smem_desc_join(hi, o_lo + 2 * k16), ptx::smem_desc_join(hi, w_lo + 2 * k16), idw0,
                           acc);
          if (cn1 > 0)
            ptx::mma_bf16_ss(tmem_base + 192, ptx::smem_desc_join(hi, o_lo + 2 * k16),
                             ptx::smem_desc_join(hi, w_lo + ((192 * 128) >> 4) + 2 * k16), idw1, acc);
          acc = 1u;
        }
        ptx::mma_commit(vw_empty + 8u * s);
      }
      ptx::mma_commit(y_full);
    }
  } else {
    // ===== softmax / conversion / epilogue warps =====
    const int et = threadIdx.x - 64;            // 0..255
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;           // column half
    const int row = q * 32 + lane;              // query row of the tile = TMEM lane
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* p_smem = smem_gen + p.off_p;
    // ---- phase 2: softmax over the keys ----
    const int u16 = p.lp >> 4;                  // 16-column units of S
    const int u_lo = half == 0 ? 0 : (u16 + 1) / 2, u_hi = half == 0 ? (u16 + 1) / 2 : u16;
    ptx::mbar_wait(s_full, 0);
    ptx::tcgen05_fence_after();
    float m = -INFINITY;
    for (int u = u_lo; u < u_hi; ++u) {
      uint32_t r[16];
      ptx::tmem_ld_x16(t_row + u * 16, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (u * 16 + i < p.L) m = fmaxf(m, __uint_as_float(r[i]));
    }
    xch[half * kAtQ + row] = m;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    m = fmaxf(xch[row], xch[kAtQ + row]);
    asm volatile("bar.sync 1, 256;" ::: "memory");      // everyone has read the maxima before the sums overwrite them
    const float mb = m * p.scale_log2e;
    float sum = 0.f;
    for (int u = u_lo; u < u_hi; ++u) {
      uint32_t r[16];
      ptx::tmem_ld_x16(t_row + u * 16, r);
      ptx::tmem_ld_wait();
      float e[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v = (u * 16 + i < p.L) ? exp2f(fmaf(__uint_as_float(r[i]), p.scale_log2e, -mb)) : 0.f;
        const float vr = __bfloat162float(__float2bfloat16_rn(v));
        e[i] = vr;
        sum += vr;
      }
      const bf16x8 lo8 = pack8(e), hi8 = pack8(e + 8);
      *reinterpret_cast<uint4*>(p_smem + kmajor_unit_off(row, u * 16)) = *reinterpret_cast<const uint4*>(&lo8);
      *reinterpret_cast<uint4*>(p_smem + kmajor_unit_off(row, u * 16 + 8)) = *reinterpret_cast<const uint4*>(&hi8);
    }
    xch[half * kAtQ + row] = sum;
    ptx::tcgen05_fence_before();
    ptx::fence_proxy_async_smem();              // P (generic-proxy stores) -> visible to the tensor core
    ptx::mbar_arrive(p_ready);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv = 1.f / (xch[row] + xch[kAtQ + row]);
    // ---- phase 4: O / sum -> bf16 K-major operand (overwrites P: the P V MMAs have completed when o_full fires) ----
    const int c16 = (p.C + 15) >> 4;
    const int c_lo = half == 0 ? 0 : (c16 + 1) / 2, c_hi = half == 0 ? (c16 + 1) / 2 : c16;
    ptx::mbar_wait(o_full, 0);
    ptx::tcgen05_fence_after();
    for (int u = c_lo; u < c_hi; ++u) {
      uint32_t r[16];
      ptx::tmem_ld_x16(t_row + u * 16, r);
      ptx::tmem_ld_wait();
      float e[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) e[i] = __uint_as_float(r[i]) * inv;
      const bf16x8 lo8 = pack8(e), hi8 = pack8(e + 8);
      *reinterpret_cast<uint4*>(p_smem + kmajor_unit_off(row, u * 16)) = *reinterpret_cast<const uint4*>(&lo8);
      *reinterpret_cast<uint4*>(p_smem + kmajor_unit_off(row, u * 16 + 8)) = *reinterpret_cast<const uint4*>(&hi8);
    }
    ptx::tcgen05_fence_before();
    ptx::fence_proxy_async_smem();
    ptx::mbar_arrive(o_ready);
    // ---- phase 6: out = (Y + bo + x) * scale ----
    const int qrow = q0 + row;
    const bool valid = qrow < p.L;
    const long long pix = (long long)b * p.L + qrow;
    const __nv_bfloat16* res_row = p.res + pix * p.res_pitch;
    __nv_bfloat16* out_row = p.out + pix * p.out_pitch;
    ptx::mbar_wait(y_full, 0);
    ptx::tcgen05_fence_after();
    for (int u = c_lo; u < c_hi; ++u) {
      uint32_t r[16];
      ptx::tmem_ld_x16(t_row + u * 16, r);
      uint4 x0 = make_uint4(0, 0, 0, 0), x1 = x0;
      const int n = u * 16;
      if (valid) {
        x0 = __ldg(reinterpret_cast<const uint4*>(res_row + n));
        x1 = __ldg(reinterpret_cast<const uint4*>(res_row + n) + 1);
      }
      ptx::tmem_ld_wait();
      if (valid) {
        float f[16], v[16];
        bf16x8 r0, r1;
        *reinterpret_cast<uint4*>(&r0) = x0;
        *reinterpret_cast<uint4*>(&r1) = x1;
        unpack8(r0, f);
        unpack8(r1, f + 8);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (__uint_as_float(r[i]) + __ldg(p.bo + n + i) + f[i]) * p.out_scale;
        const bf16x8 o0 = pack8(v), o1 = pack8(v + 8);
        reinterpret_cast<uint4*>(out_row + n)[0] = *reinterpret_cast<const uint4*>(&o0);
        reinterpret_cast<uint4*>(out_row + n)[1] = *reinterpret_cast<const uint4*>(&o1);
      }
    }
    (void)et;
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

struct AttnLayout {
  uint32_t a_slot, off_p, off_vw, vw_slot, bar_off;
  int a_slots, vw_slots, nc, nk, kbox, lp;
  size_t smem;
};

static bool attn_layout(int L, int C, AttnLayout* lay) {
  if (L < 1 || L > 512 || C < 16 || C > 320 || C % 16 != 0) return false;
  lay->nc = ceil_div(C, 64);
  lay->nk = ceil_div(L, 64);
  lay->kbox = ceil_div(L, 128);
  lay->lp = ceil_div(L, 16) * 16;
  if (lay->lp > 512) return false;
  lay->a_slot = (uint32_t)(kAtQ * 128 + lay->kbox * 128 * 128);
  lay->a_slots = lay->nc >= 2 ? 2 : 1;
  const uint32_t region_a = lay->a_slots * lay->a_slot;
  const uint32_t p_bytes = (uint32_t)(std::max(lay->nk, lay->nc) * kAtQ * 128);     // P blocks, later O blocks
  lay->off_p = 0;
  lay->off_vw = std::max(region_a, p_bytes);
  lay->vw_slot = (uint32_t)(lay->nc * 64 * 128);
  const size_t tail = 176 + 2 * kAtQ * sizeof(float) + 64;
  int slots = kAtMaxRing;
  while (slots >= 2 && 1024 + lay->off_vw + (size_t)slots * lay->vw_slot + tail > 227 * 1024) --slots;
  if (slots < 2) return false;
  lay->vw_slots = std::min(slots, lay->nk + lay->nc);
  if (lay->vw_slots < 1) lay->vw_slots = 1;
  lay->bar_off = lay->off_vw + lay->vw_slots * lay->vw_slot;
  lay->smem = 1024 + lay->bar_off + tail;
  return lay->smem <= 227 * 1024;
}

}  // namespace csd

extern "C" {

int csd_attn_core_supported(int L, int C) {
  csd::AttnLayout lay;
  return csd::attn_layout(L, C, &lay) ? 1 : 0;
}

int csd_attn_core_bf16(const void* qkv, int qkv_pitch, const void* wo, int wo_pitch, int wo_rows, const float* bo,
                       const void* res, int res_pitch, void* out, int out_pitch, int batch, int L, int C,
                       float out_scale, csd_stream_t stream_) {
  using namespace csd;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CSD_REQUIRE(qkv && wo && bo && res && out && batch >= 1 && batch <= 65535, "attn_core: bad arguments");
  AttnLayout lay;
  if (!attn_layout(L, C, &lay))
    return set_error(CSD_ERR_UNSUPPORTED, "attn_core: L=%d C=%d does not fit the fused kernel (ask csd_attn_core_supported)",
                     L, C);
  CSD_REQUIRE(qkv_pitch >= 3 * C && qkv_pitch % 8 == 0 && C % 8 == 0, "attn_core: qkv pitch %d for C=%d", qkv_pitch, C);
  CSD_REQUIRE(wo_pitch >= C && wo_pitch % 8 == 0 && wo_rows >= C, "attn_core: projection weights [%d, %d] for C=%d", wo_rows,
              wo_pitch, C);
  CSD_REQUIRE(res_pitch % 8 == 0 && out_pitch % 8 == 0 && res_pitch >= C && out_pitch >= C, "attn_core: residual / output pitch");
  CUtensorMap mq, mk, mv, mw;
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
  uint64_t dims[3] = {(uint64_t)C, (uint64_t)L, (uint64_t)batch};
  uint64_t strides[2] = {(uint64_t)qkv_pitch * 2, (uint64_t)qkv_pitch * 2 * (uint64_t)L};
  uint32_t box_qk[3] = {64, 128, 1}, box_v[3] = {64, 64, 1};
  int st = encode_tensor_map(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box_qk, TMA_SW_128);
  if (st != CSD_OK) return st;
  st = encode_tensor_map(&mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base + C, dims, strides, box_qk, TMA_SW_128);
  if (st != CSD_OK) return st;
  st = encode_tensor_map(&mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base + 2 * C, dims, strides, box_v, TMA_SW_128);
  if (st != CSD_OK) return st;
  uint64_t wdims[3] = {(uint64_t)wo_pitch, (uint64_t)wo_rows, 1};
  uint64_t wstr[2] = {(uint64_t)wo_pitch * 2, (uint64_t)wo_pitch * 2 * (uint64_t)wo_rows};
  st = encode_tensor_map(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wo, wdims, wstr, box_v, TMA_SW_128);
  if (st != CSD_OK) return st;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.L = L; p.C = C; p.batch = batch;
  p.nc = lay.nc; p.nk = lay.nk; p.kbox = lay.kbox; p.lp = lay.lp;
  p.a_slot = lay.a_slot; p.off_p = lay.off_p; p.off_vw = lay.off_vw; p.vw_slot = lay.vw_slot; p.bar_off = lay.bar_off;
  p.a_slots = lay.a_slots; p.vw_slots = lay.vw_slots;
  p.scale_log2e = 1.4426950408889634f / sqrtf((float)C);
  p.out_scale = out_scale;
  p.bo = bo;
  p.res = static_cast<const __nv_bfloat16*>(res); p.res_pitch = res_pitch;
  p.out = static_cast<__nv_bfloat16*>(out); p.out_pitch = out_pitch;
  static bool attr_set = false;
  if (!attr_set) {
    CSD_CUDA(cudaFuncSetAttribute(attn_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  attn_core_kernel<<<dim3((unsigned)ceil_div(L, kAtQ), (unsigned)batch), kAtThreads, lay.smem, stream>>>(mq, mk, mv, mw, p);
  CSD_LAUNCH_CHECK("attn_core_kernel");
  return CSD_OK;
}

}  // extern "C"
