// GroupNorm statistics / apply(+SiLU) over NHWC bf16 activations, layout conversion at the network
// boundary, attention softmax, and the time-embedding MLP.
//
// References: nn.GroupNorm(min(C//4,32), C, eps=1e-6) + nn.SiLU call sites in
// models/layerspp.py:67,219,231,242,264 and models/ncsnpp.py:331-352; F.softmax in
// models/layerspp.py:82-85; models/layers.py:524-538 / layerspp.py:32-41 / ncsnpp.py:242-260 (temb).
// The statistics are taken over the channel-concatenation of up to two tensors so that
// `torch.cat([h, hs.pop()], dim=1)` (models/ncsnpp.py:325) is never materialised on its own: the
// apply pass writes the normalised, activated concatenation directly.
#include "common.cuh"
#include "../../include/csd_b200.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace csd {

template <typename VT>
struct GnParamsT {
  const VT* src0;
  const VT* src1;
  int v0, v1;          // 8-channel vectors taken from each source
  int pv0, pv1;        // pitch of each source in vectors
  int batch, hw, groups, cpg;
  float* sums0;        // per-channel (sum, sum of squares): [batch, c0, 2]
  float* sums1;        //                                     [batch, c1, 2]
  const float* gamma;
  const float* beta;
  VT* out;
  int out_pv;
  float eps;
  int silu;
  int slabs;
};
using GnParams = GnParamsT<bf16x8>;

template <typename VT>
__device__ __forceinline__ VT gn_load(const GnParamsT<VT>& p, int b, long long pix, int v) {
  if (v < p.v0) return p.src0[((long long)b * p.hw + pix) * p.pv0 + v];
  return p.src1[((long long)b * p.hw + pix) * p.pv1 + (v - p.v0)];
}

// Per-channel statistics of ONE tensor, deterministic (no atomics: the reference is bitwise reproducible under a seed,
// SURVEY.md §8c). grid = (V / Vs, batch): a CTA owns Vs channel vectors of one image and walks ALL its pixels with
// ppb = 256 / Vs pixel lanes (4 independent vector loads in flight per thread); the lanes are reduced by a
// fixed-order tree in shared memory and the sums are stored, not accumulated. Each tensor's sums are computed once and
// shared by every GroupNorm that reads the tensor (the down-path activations are normalised twice: by the next
// block and, through the skip concatenation, by the up path).
template <typename VT>
__global__ void __launch_bounds__(256) gn_chan_stats_kernel(GnParamsT<VT> p, int Vs) {
  extern __shared__ float sm[];     // [ppb][16 * Vs] (sum, sumsq) interleaved per channel
  const int b = blockIdx.y;
  const int ppb = blockDim.x / Vs;
  const int v = threadIdx.x % Vs, pp = threadIdx.x / Vs;
  const int gv = blockIdx.x * Vs + v;
  const int n2 = 16 * Vs;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (pp < ppb && gv < p.v0) {
    long long pix = pp;
    const long long hi = p.hw;
    for (; pix + 3LL * ppb < hi; pix += 4LL * ppb) {   // 4 independent vector loads in flight
      VT v4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v4[u] = gn_load(p, b, pix + (long long)u * ppb, gv);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v4[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += f[i];
          q[i] = fmaf(f[i], f[i], q[i]);
        }
      }
    }
    for (; pix < hi; pix += ppb) {
      float f[8];
      unpack8(gn_load(p, b, pix, gv), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] = fmaf(f[i], f[i], q[i]);
      }
    }
  }
  if (pp < ppb) {
    float* row = sm + (size_t)pp * n2 + v * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      row[2 * i] = s[i];
      row[2 * i + 1] = q[i];
    }
  }
  __syncthreads();
  // fixed-order reduction over the pixel lanes: lanes are folded in halves, then one thread per slot stores
  for (int c2 = threadIdx.x; c2 < n2; c2 += blockDim.x) {
    const int ch = blockIdx.x * Vs * 8 + (c2 >> 1);
    if (ch >= p.v0 * 8) continue;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int r = 0;
    for (; r + 3 < ppb; r += 4) {
      a0 += sm[(size_t)r * n2 + c2];
      a1 += sm[(size_t)(r + 1) * n2 + c2];
      a2 += sm[(size_t)(r + 2) * n2 + c2];
      a3 += sm[(size_t)(r + 3) * n2 + c2];
    }
    for (; r < ppb; ++r) a0 += sm[(size_t)r * n2 + c2];
    p.sums0[((long long)b * p.v0 * 8 + ch) * 2 + (c2 & 1)] = (a0 + a1) + (a2 + a3);
  }
}

// chan_sums[b, c] = sum over the image's pixel tiles of the per-tile partial sums written by the
// transposed convolution epilogue (deterministic: no atomics).
__global__ void __launch_bounds__(256)
gn_finalize_partials_kernel(const float* __restrict__ partials, float* __restrict__ chan_sums, int tiles_per_img,
                            int c2 /* channels * 2 */) {
  // block (32, 8): 32 consecutive (channel, sum|sumsq) slots x 8 tile lanes; fixed-order tree -> deterministic
  __shared__ float red[8][33];
  const int b = blockIdx.y;
  const int i = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (i < c2) {
    const float* p0 = partials + (long long)b * tiles_per_img * c2 + i;
    for (int t = threadIdx.y; t < tiles_per_img; t += 8) acc += p0[(long long)t * c2];
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && i < c2) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += red[k][threadIdx.x];
    chan_sums[(long long)b * c2 + i] = a;
  }
}

// Normalise + affine (+SiLU) over the channel concatenation of up to two tensors; the group statistics
// are assembled from the tensors' per-channel sums. Dynamic smem: 2*C + 2*G floats.
template <typename VT>
__global__ void gn_apply_kernel(GnParamsT<VT> p) {
  extern __shared__ float sm[];
  const int V = p.v0 + p.v1, C = V * 8, C0 = p.v0 * 8;
  float* scale = sm;
  float* shift = sm + C;
  float* gmean = sm + 2 * C;
  float* grstd = gmean + p.groups;
  const int b = blockIdx.x / p.slabs, slab = blockIdx.x % p.slabs;
  const float inv_n = 1.f / ((float)p.hw * (float)p.cpg);
  for (int g = threadIdx.x; g < p.groups; g += blockDim.x) {
    float su = 0.f, sq = 0.f;
    for (int i = 0; i < p.cpg; ++i) {
      const int c = g * p.cpg + i;
      const float* sp = (c < C0) ? p.sums0 + ((long long)b * C0 + c) * 2
                                 : p.sums1 + ((long long)b * (C - C0) + (c - C0)) * 2;
      su += sp[0];
      sq += sp[1];
    }
    const float mean = su * inv_n;
    const float var = fmaxf(sq * inv_n - mean * mean, 0.f);
    gmean[g] = mean;
    grstd[g] = rsqrtf(var + p.eps);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / p.cpg;
    const float sc = grstd[g] * p.gamma[c];
    scale[c] = sc;
    shift[c] = p.beta[c] - gmean[g] * sc;
  }
  __syncthreads();
  const int ppb = blockDim.x / V;
  const int v = threadIdx.x % V, pp = threadIdx.x / V;
  if (pp >= ppb) return;
  const long long chunk = ceil_div_ll(p.hw, p.slabs);
  const long long lo = slab * chunk, hi = min((long long)p.hw, lo + chunk);
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = scale[v * 8 + i];
    sh[i] = shift[v * 8 + i];
  }
  long long pix = lo + pp;
  for (; pix + 3LL * ppb < hi; pix += 4LL * ppb) {
    VT v4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v4[u] = gn_load(p, b, pix + (long long)u * ppb, v);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(v4[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = fmaf(f[i], sc[i], sh[i]);
        y = (p.silu & 1) ? silu_act<VT>(y) : y;
        f[i] = (sizeof(VT) == 32 && (p.silu & 2)) ? round_tf32(y) : y;
      }
      p.out[((long long)b * p.hw + pix + (long long)u * ppb) * p.out_pv + v] = pack8_as<VT>(f);
    }
  }
  for (; pix < hi; pix += ppb) {
    float f[8];
    unpack8(gn_load(p, b, pix, v), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = fmaf(f[i], sc[i], sh[i]);
      y = (p.silu & 1) ? silu_act<VT>(y) : y;
      f[i] = (sizeof(VT) == 32 && (p.silu & 2)) ? round_tf32(y) : y;
    }
    p.out[((long long)b * p.hw + pix) * p.out_pv + v] = pack8_as<VT>(f);
  }
}

// (scale, shift) per (image, channel) of a GroupNorm over the channel concatenation of up to two tensors:
// coef[b, c] = (rstd_g * gamma_c, beta_c - mean_g * rstd_g * gamma_c), written per source tensor so that a
// convolution segment can look its channels up directly. Consumed by the fused GroupNorm+SiLU prologue of
// the transposed convolution kernel (conv_gemm.cu), which replaces gn_apply for those layers.
// grid = batch, block = 256. cpg <= 18 for every NCSN++ width, so the per-channel group loop is short.
__global__ void __launch_bounds__(256)
gn_coeffs_kernel(const float* __restrict__ sums0, int c0, const float* __restrict__ sums1, int c1,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float2* __restrict__ coef0,
                 float2* __restrict__ coef1, int hw, int cpg, float eps) {
  const int b = blockIdx.x, C = c0 + c1;
  const float inv_n = 1.f / ((float)hw * (float)cpg);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g0 = (c / cpg) * cpg;
    float su = 0.f, sq = 0.f;
    for (int i = 0; i < cpg; ++i) {
      const int cc = g0 + i;
      const float* sp = (cc < c0) ? sums0 + ((long long)b * c0 + cc) * 2 : sums1 + ((long long)b * c1 + (cc - c0)) * 2;
      su += sp[0];
      sq += sp[1];
    }
    const float mean = su * inv_n;
    const float var = fmaxf(sq * inv_n - mean * mean, 0.f);
    const float sc = rsqrtf(var + eps) * gamma[c];
    const float2 v = make_float2(sc, beta[c] - mean * sc);
    if (c < c0) coef0[(long long)b * c0 + c] = v;
    else coef1[(long long)b * c1 + (c - c0)] = v;
  }
}

// gn_coeffs with the finalize pass folded in: a source may arrive as the per-tile partial sums of the transposed
// convolution's epilogue ([batch * tiles, c, 2]); the CTA of image b reduces them in a fixed order (lanes of tiles,
// then lanes), writes the channel sums for later consumers (skip connections) and goes on to the coefficients.
// One launch instead of gn_finalize_partials + gn_coeffs. grid = (batch, slices of whole groups), block = 256.
struct GnCoefSrc {
  const float* sums;      // [batch, c, 2] or null
  const float* partials;  // [batch * tiles, c, 2] or null
  float* sums_out;        // where the reduced partials go (may be null)
  int tiles, c;
};

__global__ void __launch_bounds__(256)
gn_coeffs_partials_kernel(GnCoefSrc s0, GnCoefSrc s1, const float* __restrict__ gamma, const float* __restrict__ beta,
                          float2* __restrict__ coef0, float2* __restrict__ coef1, int hw, int cpg, float eps,
                          int slice /* channels per CTA, a multiple of cpg */) {
  // grid = (batch, slices): CTA (b, y) owns channels [y * slice, (y + 1) * slice) of the concatenation
  extern __shared__ float sm[];
  const int b = blockIdx.x, c0 = s0.c, c1 = s1.c, C = c0 + c1;
  const int cs = blockIdx.y * slice, ce = min(C, cs + slice);
  if (cs >= ce) return;
  const int n2 = 2 * (ce - cs);
  const int lanes = max(1, min(32, (int)blockDim.x / n2));
  float* sums = sm;              // [n2]
  float* red = sm + n2;          // [lanes][n2]
  for (int idx = threadIdx.x; idx < lanes * n2; idx += blockDim.x) {
    const int i = idx % n2, g = idx / n2;
    const int cc = cs + (i >> 1), comp = i & 1;
    const bool first = cc < c0;
    const GnCoefSrc& s = first ? s0 : s1;
    const int lc = first ? cc : cc - c0;
    float acc = 0.f;
    if (s.partials != nullptr) {
      const long long stride = 2LL * s.c;
      const float* p0 = s.partials + (long long)b * s.tiles * stride + 2 * lc + comp;
      float a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int t = g;
      for (; t + 3 * lanes < s.tiles; t += 4 * lanes) {     // 4 independent loads in flight per thread
        acc += p0[(long long)t * stride];
        a1 += p0[(long long)(t + lanes) * stride];
        a2 += p0[(long long)(t + 2 * lanes) * stride];
        a3 += p0[(long long)(t + 3 * lanes) * stride];
      }
      for (; t < s.tiles; t += lanes) acc += p0[(long long)t * stride];
      acc = (acc + a1) + (a2 + a3);
    } else if (g == 0) {
      acc = s.sums[((long long)b * s.c + lc) * 2 + comp];
    }
    red[idx] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    float a = 0.f;
    for (int g = 0; g < lanes; ++g) a += red[g * n2 + i];
    sums[i] = a;
    const int cc = cs + (i >> 1);
    const bool first = cc < c0;
    const GnCoefSrc& s = first ? s0 : s1;
    if (s.partials != nullptr && s.sums_out != nullptr)
      s.sums_out[((long long)b * s.c + (first ? cc : cc - c0)) * 2 + (i & 1)] = a;
  }
  __syncthreads();
  const float inv_n = 1.f / ((float)hw * (float)cpg);
  for (int c = cs + threadIdx.x; c < ce; c += blockDim.x) {
    const int g0 = ((c - cs) / cpg) * cpg;
    float su = 0.f, sq = 0.f;
    for (int i = 0; i < cpg; ++i) {
      su += sums[2 * (g0 + i)];
      sq += sums[2 * (g0 + i) + 1];
    }
    const float mean = su * inv_n;
    const float var = fmaxf(sq * inv_n - mean * mean, 0.f);
    const float sc = rsqrtf(var + eps) * gamma[c];
    const float2 v = make_float2(sc, beta[c] - mean * sc);
    if (c < c0) coef0[(long long)b * c0 + c] = v;
    else coef1[(long long)b * c1 + (c - c0)] = v;
  }
}

template <typename VT>
static int gn_fill(GnParamsT<VT>& p, const void* src0, int c0, int pitch0, const void* src1, int c1, int pitch1, int batch,
                   int hw, int groups, int* threads, size_t* smem) {
  CSD_REQUIRE(src0 != nullptr && c0 >= 8 && c0 % 8 == 0 && pitch0 % 8 == 0 && c0 <= pitch0,
              "groupnorm: source 0 needs >= 8 channels in multiples of 8 (c=%d pitch=%d)", c0, pitch0);
  CSD_REQUIRE(src1 == nullptr || (c1 >= 8 && c1 % 8 == 0 && pitch1 % 8 == 0 && c1 <= pitch1),
              "groupnorm: source 1 channels must be a multiple of 8 (c=%d pitch=%d)", c1, pitch1);
  if (src1 == nullptr) c1 = 0;
  const int C = c0 + c1;
  CSD_REQUIRE(groups >= 1 && C % groups == 0, "groupnorm: %d channels not divisible by %d groups", C, groups);
  const int V = C / 8;
  CSD_REQUIRE(V <= 1024, "groupnorm: %d channels exceed the 8192 supported", C);
  p.src0 = static_cast<const VT*>(src0);
  p.src1 = static_cast<const VT*>(src1);
  p.v0 = c0 / 8; p.v1 = c1 / 8; p.pv0 = pitch0 / 8; p.pv1 = src1 ? pitch1 / 8 : 0;
  p.batch = batch; p.hw = hw; p.groups = groups; p.cpg = C / groups;
  const int ppb = std::max(1, 256 / V);
  *threads = V * ppb;
  *smem = sizeof(float) * (2 * C + 2 * groups);
  int slabs = ceil_div(num_sms() * 16, batch);
  const int max_slabs = std::max(1, hw / (ppb * 8));
  p.slabs = std::max(1, std::min(slabs, max_slabs));
  return CSD_OK;
}

// ---- one-launch GroupNorm for the small levels ---------------------------------------------------------
// At <= 20 px the producer of a tensor does not deliver channel sums for free (only the transposed convolution
// does) and a whole image is a few KB, so statistics -> apply as separate launches is three latency-bound
// kernels per GroupNorm. Here one CTA owns (image b, a slice of Vs 8-channel vectors that covers whole groups):
// it reads its [hw, Vs] sub-tensor ONCE into registers, reduces per-channel sums in shared memory (fixed-order
// tree, deterministic), forms the group statistics of its own groups, and writes the normalised (+SiLU)
// result from the registers. grid = (slices, batch); block = Vs * ppb threads; hw <= R * ppb.
struct GnFusedPlan {
  int vs;        // vectors per slice
  int slices;
  int threads;
  int ppb;       // pixel lanes per CTA
  int r;         // register-cached vectors per thread (template instance)
  size_t smem;
};

// Makes the compiler forget what it knows about the words of a register-cached vector: without it the unpacked fp32
// values of the statistics pass are kept alive across the barriers for the apply pass (2-4x the registers of the packed
// form, spilled under the 2-CTAs-per-SM bound) instead of being unpacked again.
template <typename VT> __device__ __forceinline__ void keep_packed(VT& v) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(VT) / 4); ++i) asm volatile("" : "+r"(w[i]));
}

template <int R, typename VT>
__global__ void __launch_bounds__(512, (R <= 8 && sizeof(VT) == 16) ? 2 : 1)
gn_fused_kernel(GnParamsT<VT> p, int Vs) {
  extern __shared__ float sm[];
  const int b = blockIdx.y;
  const int ppb = blockDim.x / Vs;
  const int v = threadIdx.x % Vs, pp = threadIdx.x / Vs;
  const int gv = blockIdx.x * Vs + v;       // vector index in the channel concatenation
  const int n2 = Vs * 16;                   // (sum, sumsq) slots of the slice
  float* part = sm;                         // [ppb][n2]
  float* red = sm + (size_t)ppb * n2;       // [8][n2]
  float* chs = red + 8 * n2;                // [n2]
  float* coef = chs + n2;                   // [n2] = (scale, shift) per channel of the slice

  VT cache[R];
  {
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int pix = pp + u * ppb;
      if (pix < p.hw) cache[u] = gn_load(p, b, pix, gv);
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int pix = pp + u * ppb;
      if (pix < p.hw) {
        float f[8];
        unpack8(cache[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += f[i];
          q[i] = fmaf(f[i], f[i], q[i]);
        }
      }
    }
    float4* row = reinterpret_cast<float4*>(part + (size_t)pp * n2 + v * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) row[i] = make_float4(s[2 * i], q[2 * i], s[2 * i + 1], q[2 * i + 1]);
#pragma unroll
    for (int u = 0; u < R; ++u) keep_packed(cache[u]);
  }
  __syncthreads();
  // two-level fixed-order reduction over the pixel lanes
  const int parts = min(8, max(1, (int)blockDim.x / n2));
  for (int idx = threadIdx.x; idx < parts * n2; idx += blockDim.x) {
    const int c2 = idx % n2, pt = idx / n2;
    float a = 0.f;
    for (int r = pt; r < ppb; r += parts) a += part[(size_t)r * n2 + c2];
    red[pt * n2 + c2] = a;
  }
  __syncthreads();
  for (int c2 = threadIdx.x; c2 < n2; c2 += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < parts; ++k) a += red[k * n2 + c2];
    chs[c2] = a;
  }
  __syncthreads();
  // one thread per channel of the slice: statistics of its group (every member of a group adds the same channel sums
  // in the same order), then the channel's (scale, shift)
  for (int lc = threadIdx.x; lc < Vs * 8; lc += blockDim.x) {
    const int g0 = (lc / p.cpg) * p.cpg;
    float su = 0.f, sq = 0.f;
    for (int i = 0; i < p.cpg; ++i) {
      const float2 t = *reinterpret_cast<const float2*>(chs + 2 * (g0 + i));
      su += t.x;
      sq += t.y;
    }
    const float inv_n = 1.f / ((float)p.hw * (float)p.cpg);
    const float mean = su * inv_n;
    const float var = fmaxf(sq * inv_n - mean * mean, 0.f);
    const int c = blockIdx.x * Vs * 8 + lc;
    const float a = rsqrtf(var + p.eps) * __ldg(p.gamma + c);
    *reinterpret_cast<float2*>(coef + 2 * lc) = make_float2(a, __ldg(p.beta + c) - mean * a);
  }
  __syncthreads();
  float sc[8], sh[8];
  {
    const float4* cp = reinterpret_cast<const float4*>(coef + v * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = cp[i];
      sc[2 * i] = t.x; sh[2 * i] = t.y; sc[2 * i + 1] = t.z; sh[2 * i + 1] = t.w;
    }
  }
#pragma unroll
  for (int u = 0; u < R; ++u) {
    const int pix = pp + u * ppb;
    if (pix < p.hw) {
      float f[8];
      unpack8(cache[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = fmaf(f[i], sc[i], sh[i]);
        y = (p.silu & 1) ? silu_act<VT>(y) : y;
        f[i] = (sizeof(VT) == 32 && (p.silu & 2)) ? round_tf32(y) : y;
      }
      p.out[((long long)b * p.hw + pix) * p.out_pv + gv] = pack8_as<VT>(f);
    }
  }
}

static int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

// Host-side planning (no CUDA calls): returns false when the shape does not fit the one-launch kernel.
static bool gn_fused_plan_with(int min_vs, int max_threads, int C, int hw, int cpg, int batch, int sms, GnFusedPlan* pl,
                               int elem_bytes) {
  const int V = C / 8;
  const int base = cpg / gcd_int(cpg, 8);   // lcm(cpg, 8) / 8 vectors: smallest slice made of whole groups
  if (V % base) return false;
  const int nb = V / base;
  // most slices whose rows are still >= min_vs vectors unless the tensor is narrower than that; with a big batch
  // fewer, wider slices are enough to fill the machine
  int best = 1;
  for (int d = nb; d >= 1; --d) {
    if (nb % d) continue;
    const int vs = V / d;
    if (vs >= min_vs || d == 1) { best = d; break; }
  }
  while (best > 1) {   // shrink while the grid stays >= 2 waves and the next divisor exists
    int d2 = best - 1;
    while (d2 >= 1 && nb % d2) --d2;
    if (d2 < 1 || (long long)batch * d2 < 2LL * sms) break;
    const int vs2 = V / d2;
    if (vs2 > 128 || ceil_div(hw, 512 / vs2) > 8) break;   // stay within the register-cached iteration budget
    best = d2;
  }
  const int vs = V / best;
  if (vs > 128) return false;
  int threads = 256;
  int ppb = threads / vs;
  if (ppb < 1) return false;
  if (ceil_div(hw, ppb) > 8 && max_threads >= 512) { threads = 512; ppb = threads / vs; }
  const int iters = ceil_div(hw, ppb);
  if (iters > (elem_bytes == 4 ? 8 : 16)) return false;   // register-cached vectors per thread (fp32 vectors are 8 registers)
  pl->vs = vs;
  pl->slices = best;
  pl->ppb = ppb;
  pl->threads = vs * ppb;
  pl->r = iters <= 1 ? 1 : (iters <= 2 ? 2 : (iters <= 4 ? 4 : (iters <= 8 ? 8 : 16)));
  const int n2 = vs * 16;
  pl->smem = sizeof(float) * ((size_t)ppb * n2 + 8 * n2 + n2 + n2);
  return pl->smem <= 96 * 1024;
}

static bool gn_fused_plan(int c0, int c1, int hw, int groups, int batch, int sms, GnFusedPlan* pl, int elem_bytes = 2) {
  const int C = c0 + c1;
  if (c0 < 8 || c0 % 8 || c1 % 8 || groups < 1 || C % groups || hw < 1 || batch < 1) return false;
  const int cpg = C / groups;
  int max_threads = 512;
  if (const char* e = getenv("CSD_GNF_THREADS")) max_threads = std::max(64, atoi(e));   // probe only
  if (const char* e = getenv("CSD_GNF_MINVEC"))                                         // probe only
    return gn_fused_plan_with(std::max(1, atoi(e)), max_threads, C, hw, cpg, batch, sms, pl, elem_bytes);
  // 20 px class images in bf16: rows of >= 8 vectors (128 B) and half as many, fatter CTAs measured 9.8 vs 11.8 us at
  // [64, 400, 192] (tools/gn_fused_probe.py); everything else, and whatever does not fit that way: >= 4 vectors
  if (elem_bytes == 2 && hw >= 256 && gn_fused_plan_with(8, max_threads, C, hw, cpg, batch, sms, pl, elem_bytes)) return true;
  return gn_fused_plan_with(4, max_threads, C, hw, cpg, batch, sms, pl, elem_bytes);
}

// ---- layout conversion ---------------------------------------------------------------------------
template <typename VT>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ s0, int c0, const float* __restrict__ s1, int c1, VT* __restrict__ out,
                    int cvec, int batch, long long hw, float scale, float shift) {
  const long long total = (long long)batch * hw * cvec;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    // pixel index fastest inside a channel vector so the fp32 plane reads coalesce
    const long long pix = idx % hw;
    const int cv = (int)((idx / hw) % cvec);
    const int b = (int)(idx / (hw * cvec));
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = cv * 8 + i;
      float v = 0.f;
      if (c < c0) v = fmaf(__ldg(s0 + ((long long)b * c0 + c) * hw + pix), scale, shift);
      else if (c < c0 + c1) v = fmaf(__ldg(s1 + ((long long)b * c1 + (c - c0)) * hw + pix), scale, shift);
      f[i] = sizeof(VT) == 32 ? round_tf32(v) : v;    // fp32 plan: the network input only feeds tensor-core operands
    }
    out[((long long)b * hw + pix) * cvec + cv] = pack8_as<VT>(f);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const T* __restrict__ src, int pitch, int c_off, int c_cnt, float* __restrict__ dst,
                    int batch, long long hw, const float* __restrict__ row_scale) {
  const long long total = (long long)batch * c_cnt * hw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long pix = idx % hw;
    const int c = (int)((idx / hw) % c_cnt);
    const int b = (int)(idx / (hw * c_cnt));
    float v = to_f32(src[((long long)b * hw + pix) * pitch + c_off + c]);
    if (row_scale != nullptr) v *= __ldg(row_scale + b);
    dst[idx] = v;
  }
}

// ---- softmax ---------------------------------------------------------------------------------------
// One warp per row; rows are short (<= 1024 keys), so the three passes hit L1.
template <typename T> __device__ __forceinline__ float store_round(float v) { return v; }
template <> __device__ __forceinline__ float store_round<float>(float v) { return round_tf32(v); }   // P feeds the PV GEMM

template <typename T>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ logits, int in_pitch, T* __restrict__ probs, int out_pitch,
                    long long rows, int cols, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* in = logits + row * in_pitch;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, in[c] * scale);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += __expf(in[c] * scale - m);
  s = warp_sum(s);
  const float inv = 1.f / s;
  T* out = probs + row * out_pitch;
  for (int c = lane; c < out_pitch; c += 32)
    out[c] = from_f32<T>(c < cols ? store_round<T>(__expf(in[c] * scale - m) * inv) : 0.f);
}

// ---- time embedding ----------------------------------------------------------------------------------
// One CTA per batch row. emb -> Linear -> SiLU -> Linear -> SiLU (the SiLU every block applies to temb
// before its Dense_0, models/layerspp.py:263, is hoisted here).
__global__ void __launch_bounds__(512)
time_embedding_kernel(const float* __restrict__ labels, int nf, int embedding_type, const float* __restrict__ fourier_w,
                      const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
                      const float* __restrict__ b1, float* __restrict__ act_temb) {
  extern __shared__ float sm[];
  const int embed = embedding_type == 1 ? 2 * nf : nf;
  const int hid = 4 * nf;
  float* emb = sm;            // [embed]
  float* h0 = sm + embed;     // [hid]
  const int b = blockIdx.x;
  const float t = labels[b];
  if (embedding_type == 1) {
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
      const float a = t * fourier_w[i] * 2.f * 3.14159265358979323846f;
      emb[i] = sinf(a);
      emb[nf + i] = cosf(a);
    }
  } else {
    const int half = nf / 2;
    const float c = logf(10000.f) / (float)(half - 1);
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
      const float a = t * expf(-c * (float)i);
      emb[i] = sinf(a);
      emb[half + i] = cosf(a);
    }
    if ((nf & 1) && threadIdx.x == 0) emb[nf - 1] = 0.f;
  }
  __syncthreads();
  // one warp per output: lanes stride the input dimension, so the weight row reads coalesce
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int j = warp; j < hid; j += nwarps) {
    const float* wr = w0 + (long long)j * embed;
    float acc = 0.f;
    for (int k = lane; k < embed; k += 32) acc = fmaf(__ldg(wr + k), emb[k], acc);
    acc = warp_sum(acc) + b0[j];
    if (lane == 0) h0[j] = acc / (1.f + expf(-acc));  // SiLU, full-precision exp
  }
  __syncthreads();
  for (int j = warp; j < hid; j += nwarps) {
    const float* wr = w1 + (long long)j * hid;
    float acc = 0.f;
    for (int k = lane; k < hid; k += 32) acc = fmaf(__ldg(wr + k), h0[k], acc);
    acc = warp_sum(acc) + b1[j];
    if (lane == 0) act_temb[(long long)b * hid + j] = acc / (1.f + expf(-acc));
  }
}

// out[b, j] = bias[j] + sum_k w[j, k] act[b, k]: every block's Dense_0 projection as ONE small fp32 GEMM.
// CTA tile = 64 batch rows x 64 outputs, K in chunks of 32 staged (transposed) in shared memory; each of the 256
// threads owns a 4 x 4 register tile. Both operands are read with coalesced 128-byte rows.
constexpr int kDenseTile = 64, kDenseK = 32;
__global__ void __launch_bounds__(256)
dense_rows_kernel(const float* __restrict__ act, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, int batch, int in_dim, int total_out) {
  __shared__ float As[kDenseK][kDenseTile + 4];   // [k][batch row]
  __shared__ float Ws[kDenseK][kDenseTile + 4];   // [k][output]
  const int j0 = blockIdx.x * kDenseTile, b0 = blockIdx.y * kDenseTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // outputs tx*4.., batch rows ty*4..
  const int lk = threadIdx.x & 31, lr = threadIdx.x >> 5;   // loader: k lane, row group (8 groups x 8 rows)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < in_dim; k0 += kDenseK) {
    const int k = k0 + lk;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = lr * 8 + r;
      const int bb = b0 + row, jj = j0 + row;
      As[lk][row] = (k < in_dim && bb < batch) ? __ldg(act + (long long)bb * in_dim + k) : 0.f;
      Ws[lk][row] = (k < in_dim && jj < total_out) ? __ldg(w + (long long)jj * in_dim + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kDenseK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int bb = b0 + ty * 4 + i;
    if (bb >= batch) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jj = j0 + tx * 4 + j;
      if (jj < total_out) out[(long long)bb * total_out + jj] = acc[i][j] + (bias != nullptr ? __ldg(bias + jj) : 0.f);
    }
  }
}

// ---- host launchers shared by the bf16 / fp32-activation entry points ---------------------------------------
template <typename VT>
static int gn_chan_stats_launch(const void* src, int c, int pitch, float* chan_sums, int batch, int hw, cudaStream_t stream) {
  CSD_REQUIRE(chan_sums != nullptr && batch >= 1 && hw >= 1, "gn_chan_stats: bad arguments");
  GnParamsT<VT> p;
  memset(&p, 0, sizeof(p));
  int threads;
  size_t smem;
  int st = gn_fill(p, src, c, pitch, nullptr, 0, 0, batch, hw, 1, &threads, &smem);
  if (st != CSD_OK) return st;
  p.sums0 = chan_sums;
  // Vs channel vectors per CTA: 32-byte rows per pixel at least (a full DRAM sector), more when the batch alone fills
  // the machine several times over (fewer, fatter CTAs re-use their pixel rows' sectors)
  const int V = c / 8;
  int vs = sizeof(VT) == 32 ? 1 : 2;
  while (vs * 2 <= 8 && V % (vs * 2) == 0 && (long long)batch * (V / (vs * 2)) >= 4LL * num_sms()) vs *= 2;
  if (V % vs != 0) vs = 1;
  const int ppb = 256 / vs;
  const size_t stat_smem = sizeof(float) * 16 * vs * (size_t)ppb;
  gn_chan_stats_kernel<VT><<<dim3((unsigned)(V / vs), (unsigned)batch), vs * ppb, stat_smem, stream>>>(p, vs);
  CSD_LAUNCH_CHECK("gn_chan_stats_kernel");
  return CSD_OK;
}

template <typename VT>
static int gn_apply_launch(const void* src0, int c0, int pitch0, const float* sums0, const void* src1, int c1, int pitch1,
                           const float* sums1, const float* gamma, const float* beta, void* out, int out_pitch, int batch,
                           int hw, int groups, float eps, int apply_silu, cudaStream_t stream) {
  CSD_REQUIRE(sums0 && gamma && beta && out && batch >= 1 && hw >= 1, "gn_apply: bad arguments");
  CSD_REQUIRE(src1 == nullptr || sums1 != nullptr, "gn_apply: second source without its channel sums");
  GnParamsT<VT> p;
  memset(&p, 0, sizeof(p));
  int threads;
  size_t smem;
  int st = gn_fill(p, src0, c0, pitch0, src1, c1, pitch1, batch, hw, groups, &threads, &smem);
  if (st != CSD_OK) return st;
  CSD_REQUIRE(out_pitch % 8 == 0 && out_pitch >= (p.v0 + p.v1) * 8, "gn_apply: out pitch %d too small", out_pitch);
  p.sums0 = const_cast<float*>(sums0);
  p.sums1 = const_cast<float*>(sums1);
  p.gamma = gamma; p.beta = beta;
  p.out = static_cast<VT*>(out);
  p.out_pv = out_pitch / 8;
  p.eps = eps; p.silu = apply_silu;
  gn_apply_kernel<VT><<<batch * p.slabs, threads, smem, stream>>>(p);
  CSD_LAUNCH_CHECK("gn_apply_kernel");
  return CSD_OK;
}

template <typename VT>
static int gn_fused_launch(const void* src0, int c0, int pitch0, const void* src1, int c1, int pitch1, const float* gamma,
                           const float* beta, void* out, int out_pitch, int batch, int hw, int groups, float eps,
                           int apply_silu, cudaStream_t stream) {
  CSD_REQUIRE(src0 && gamma && beta && out && batch >= 1 && hw >= 1, "gn_fused: bad arguments");
  if (src1 == nullptr) c1 = 0;
  GnFusedPlan pl;
  if (!gn_fused_plan(c0, c1, hw, groups, batch, 148, &pl, (int)sizeof(VT) / 8))
    return set_error(CSD_ERR_UNSUPPORTED, "gn_fused: shape c=%d+%d hw=%d groups=%d does not fit the one-launch kernel "
                     "(ask csd_gn_fused_supported first)", c0, c1, hw, groups);
  GnParamsT<VT> p;
  memset(&p, 0, sizeof(p));
  int threads;
  size_t smem;
  int st = gn_fill(p, src0, c0, pitch0, src1, c1, pitch1, batch, hw, groups, &threads, &smem);
  if (st != CSD_OK) return st;
  CSD_REQUIRE(out_pitch % 8 == 0 && out_pitch >= c0 + c1, "gn_fused: out pitch %d too small", out_pitch);
  p.gamma = gamma; p.beta = beta;
  p.out = static_cast<VT*>(out);
  p.out_pv = out_pitch / 8;
  p.eps = eps; p.silu = apply_silu;
  static bool attr_set = false;
  if (!attr_set) {
    CSD_CUDA(cudaFuncSetAttribute(gn_fused_kernel<1, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CSD_CUDA(cudaFuncSetAttribute(gn_fused_kernel<2, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CSD_CUDA(cudaFuncSetAttribute(gn_fused_kernel<4, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CSD_CUDA(cudaFuncSetAttribute(gn_fused_kernel<8, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    if (sizeof(VT) == 16)
      CSD_CUDA(cudaFuncSetAttribute(gn_fused_kernel<16, bf16x8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  const dim3 grid((unsigned)pl.slices, (unsigned)batch);
  switch (pl.r) {
    case 1: gn_fused_kernel<1, VT><<<grid, pl.threads, pl.smem, stream>>>(p, pl.vs); break;
    case 2: gn_fused_kernel<2, VT><<<grid, pl.threads, pl.smem, stream>>>(p, pl.vs); break;
    case 4: gn_fused_kernel<4, VT><<<grid, pl.threads, pl.smem, stream>>>(p, pl.vs); break;
    case 8: gn_fused_kernel<8, VT><<<grid, pl.threads, pl.smem, stream>>>(p, pl.vs); break;
    default:
      if constexpr (sizeof(VT) == 16) {
        gn_fused_kernel<16, VT><<<grid, pl.threads, pl.smem, stream>>>(p, pl.vs);
      } else {
        return set_error(CSD_ERR_UNSUPPORTED, "gn_fused (fp32 activations): more than 8 cached vectors per thread");
      }
      break;
  }
  CSD_LAUNCH_CHECK("gn_fused_kernel");
  return CSD_OK;
}

template <typename VT>
static int nchw_to_nhwc_launch(const float* src0, int c0, const float* src1, int c1, void* out, int c_pad, int batch, int h,
                               int w, float scale, float shift, cudaStream_t stream) {
  CSD_REQUIRE(src0 && out && c0 >= 1, "nchw_to_nhwc: bad arguments");
  if (src1 == nullptr) c1 = 0;
  CSD_REQUIRE(c_pad % 8 == 0 && c_pad >= c0 + c1, "nchw_to_nhwc: c_pad=%d must be a multiple of 8 >= %d", c_pad, c0 + c1);
  const long long total = (long long)batch * h * w * (c_pad / 8);
  const int blocks = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)num_sms() * 16);
  nchw_to_nhwc_kernel<VT><<<blocks, 256, 0, stream>>>(src0, c0, src1, c1, static_cast<VT*>(out), c_pad / 8, batch,
                                                      (long long)h * w, scale, shift);
  CSD_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return CSD_OK;
}

template <typename T>
static int nhwc_to_nchw_launch(const void* src, int c_pitch, int c_off, int c_cnt, float* dst, int batch, int h, int w,
                               const float* row_scale, cudaStream_t stream) {
  CSD_REQUIRE(src && dst && c_cnt >= 1 && c_off >= 0 && c_off + c_cnt <= c_pitch, "nhwc_to_nchw: bad channel range");
  const long long total = (long long)batch * c_cnt * h * w;
  const int blocks = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)num_sms() * 16);
  nhwc_to_nchw_kernel<T><<<blocks, 256, 0, stream>>>(static_cast<const T*>(src), c_pitch, c_off, c_cnt, dst, batch,
                                                     (long long)h * w, row_scale);
  CSD_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return CSD_OK;
}

template <typename T>
static int softmax_rows_launch(const float* logits, int in_pitch, void* probs, int out_pitch, int64_t rows, int cols,
                               float scale, cudaStream_t stream) {
  CSD_REQUIRE(logits && probs && cols >= 1 && in_pitch >= cols && out_pitch >= cols, "softmax: bad arguments");
  if (rows == 0) return CSD_OK;
  const int blocks = (int)ceil_div_ll(rows, 8);
  softmax_rows_kernel<T><<<blocks, 256, 0, stream>>>(logits, in_pitch, static_cast<T*>(probs), out_pitch, rows, cols, scale);
  CSD_LAUNCH_CHECK("softmax_rows_kernel");
  return CSD_OK;
}

}  // namespace csd

extern "C" {

int csd_gn_chan_stats_bf16(const void* src, int c, int pitch, float* chan_sums, int batch, int hw,
                           csd_stream_t stream) {
  return csd::gn_chan_stats_launch<csd::bf16x8>(src, c, pitch, chan_sums, batch, hw, static_cast<cudaStream_t>(stream));
}
int csd_gn_chan_stats_f32(const void* src, int c, int pitch, float* chan_sums, int batch, int hw, csd_stream_t stream) {
  return csd::gn_chan_stats_launch<csd::f32x8>(src, c, pitch, chan_sums, batch, hw, static_cast<cudaStream_t>(stream));
}

int csd_gn_finalize_partials_f32(const float* partials, float* chan_sums, int batch, int tiles_per_img, int c,
                                 csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(partials && chan_sums && batch >= 1 && tiles_per_img >= 1 && c >= 1, "gn_finalize_partials: bad arguments");
  dim3 grid((unsigned)ceil_div(2 * c, 32), (unsigned)batch);
  gn_finalize_partials_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(partials, chan_sums,
                                                                                          tiles_per_img, 2 * c);
  CSD_LAUNCH_CHECK("gn_finalize_partials_kernel");
  return CSD_OK;
}

int csd_gn_apply_bf16(const void* src0, int c0, int pitch0, const float* sums0, const void* src1, int c1, int pitch1,
                      const float* sums1, const float* gamma, const float* beta, void* out, int out_pitch, int batch,
                      int hw, int groups, float eps, int apply_silu, csd_stream_t stream) {
  return csd::gn_apply_launch<csd::bf16x8>(src0, c0, pitch0, sums0, src1, c1, pitch1, sums1, gamma, beta, out, out_pitch,
                                           batch, hw, groups, eps, apply_silu, static_cast<cudaStream_t>(stream));
}
int csd_gn_apply_f32(const void* src0, int c0, int pitch0, const float* sums0, const void* src1, int c1, int pitch1,
                     const float* sums1, const float* gamma, const float* beta, void* out, int out_pitch, int batch,
                     int hw, int groups, float eps, int apply_silu, csd_stream_t stream) {
  return csd::gn_apply_launch<csd::f32x8>(src0, c0, pitch0, sums0, src1, c1, pitch1, sums1, gamma, beta, out, out_pitch,
                                          batch, hw, groups, eps, apply_silu, static_cast<cudaStream_t>(stream));
}

int csd_gn_fused_supported(int c0, int c1, int hw, int groups, int batch) {
  csd::GnFusedPlan pl;
  return csd::gn_fused_plan(c0, c1, hw, groups, batch, 148, &pl) ? 1 : 0;
}
int csd_gn_fused_supported_f32(int c0, int c1, int hw, int groups, int batch) {
  csd::GnFusedPlan pl;
  return csd::gn_fused_plan(c0, c1, hw, groups, batch, 148, &pl, 4) ? 1 : 0;
}

int csd_gn_fused_bf16(const void* src0, int c0, int pitch0, const void* src1, int c1, int pitch1, const float* gamma,
                      const float* beta, void* out, int out_pitch, int batch, int hw, int groups, float eps,
                      int apply_silu, csd_stream_t stream) {
  return csd::gn_fused_launch<csd::bf16x8>(src0, c0, pitch0, src1, c1, pitch1, gamma, beta, out, out_pitch, batch, hw,
                                           groups, eps, apply_silu, static_cast<cudaStream_t>(stream));
}
int csd_gn_fused_f32(const void* src0, int c0, int pitch0, const void* src1, int c1, int pitch1, const float* gamma,
                     const float* beta, void* out, int out_pitch, int batch, int hw, int groups, float eps,
                     int apply_silu, csd_stream_t stream) {
  return csd::gn_fused_launch<csd::f32x8>(src0, c0, pitch0, src1, c1, pitch1, gamma, beta, out, out_pitch, batch, hw,
                                          groups, eps, apply_silu, static_cast<cudaStream_t>(stream));
}

int csd_gn_coeffs_f32(const float* sums0, int c0, const float* sums1, int c1, const float* gamma, const float* beta,
                      float* coef0, float* coef1, int batch, int hw, int groups, float eps, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(sums0 && gamma && beta && coef0 && batch >= 1 && hw >= 1 && c0 >= 1, "gn_coeffs: bad arguments");
  if (sums1 == nullptr) c1 = 0;
  CSD_REQUIRE(c1 == 0 || coef1 != nullptr, "gn_coeffs: second source without its coefficient output");
  CSD_REQUIRE(groups >= 1 && (c0 + c1) % groups == 0, "gn_coeffs: %d channels not divisible by %d groups", c0 + c1,
              groups);
  gn_coeffs_kernel<<<batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      sums0, c0, sums1, c1, gamma, beta, reinterpret_cast<float2*>(coef0), reinterpret_cast<float2*>(coef1), hw,
      (c0 + c1) / groups, eps);
  CSD_LAUNCH_CHECK("gn_coeffs_kernel");
  return CSD_OK;
}

int csd_gn_coeffs_partials_f32(const float* sums0, const float* partials0, int tiles0, float* sums_out0, int c0,
                               const float* sums1, const float* partials1, int tiles1, float* sums_out1, int c1,
                               const float* gamma, const float* beta, float* coef0, float* coef1, int batch, int hw,
                               int groups, float eps, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(gamma && beta && coef0 && batch >= 1 && hw >= 1 && c0 >= 1, "gn_coeffs_partials: bad arguments");
  CSD_REQUIRE((sums0 != nullptr) != (partials0 != nullptr), "gn_coeffs_partials: source 0 needs sums OR partials");
  CSD_REQUIRE(partials0 == nullptr || tiles0 >= 1, "gn_coeffs_partials: tiles0=%d", tiles0);
  if (sums1 == nullptr && partials1 == nullptr) c1 = 0;
  CSD_REQUIRE(c1 == 0 || (coef1 != nullptr && (sums1 != nullptr) != (partials1 != nullptr) &&
                          (partials1 == nullptr || tiles1 >= 1)),
              "gn_coeffs_partials: source 1 needs its coefficient output and sums OR partials");
  CSD_REQUIRE(groups >= 1 && (c0 + c1) % groups == 0, "gn_coeffs_partials: %d channels not divisible by %d groups",
              c0 + c1, groups);
  GnCoefSrc s0{sums0, partials0, sums_out0, tiles0, c0};
  GnCoefSrc s1{sums1, partials1, sums_out1, tiles1, c1};
  const int cpg = (c0 + c1) / groups;
  // Many small CTAs, each owning whole groups: the kernel is a dependent chain of L2 reads (tiles per lane), so narrower
  // slices (more lanes of tiles per channel, up to 32) shorten it; ~8 CTAs per SM keep the machine covered
  int slices = std::max(1, std::min(groups, ceil_div(8 * num_sms(), batch)));
  const int slice = ceil_div(groups, slices) * cpg;
  slices = ceil_div(c0 + c1, slice);
  const size_t smem = sizeof(float) * (2 * (size_t)slice + (size_t)std::max(256, 2 * slice) + 64);
  CSD_REQUIRE(smem <= 48 * 1024, "gn_coeffs_partials: %d channels per slice exceed the shared-memory budget", slice);
  gn_coeffs_partials_kernel<<<dim3((unsigned)batch, (unsigned)slices), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      s0, s1, gamma, beta, reinterpret_cast<float2*>(coef0), reinterpret_cast<float2*>(coef1), hw, cpg, eps, slice);
  CSD_LAUNCH_CHECK("gn_coeffs_partials_kernel");
  return CSD_OK;
}

int csd_nchw_to_nhwc_bf16(const float* src0, int c0, const float* src1, int c1, void* out, int c_pad, int batch, int h,
                          int w, float scale, float shift, csd_stream_t stream) {
  return csd::nchw_to_nhwc_launch<csd::bf16x8>(src0, c0, src1, c1, out, c_pad, batch, h, w, scale, shift,
                                               static_cast<cudaStream_t>(stream));
}
int csd_nchw_to_nhwc_f32(const float* src0, int c0, const float* src1, int c1, void* out, int c_pad, int batch, int h,
                         int w, float scale, float shift, csd_stream_t stream) {
  return csd::nchw_to_nhwc_launch<csd::f32x8>(src0, c0, src1, c1, out, c_pad, batch, h, w, scale, shift,
                                              static_cast<cudaStream_t>(stream));
}

int csd_nhwc_bf16_to_nchw(const void* src, int c_pitch, int c_off, int c_cnt, float* dst, int batch, int h, int w,
                          const float* row_scale, csd_stream_t stream) {
  return csd::nhwc_to_nchw_launch<__nv_bfloat16>(src, c_pitch, c_off, c_cnt, dst, batch, h, w, row_scale,
                                                 static_cast<cudaStream_t>(stream));
}
int csd_nhwc_f32_to_nchw(const void* src, int c_pitch, int c_off, int c_cnt, float* dst, int batch, int h, int w,
                         const float* row_scale, csd_stream_t stream) {
  return csd::nhwc_to_nchw_launch<float>(src, c_pitch, c_off, c_cnt, dst, batch, h, w, row_scale,
                                         static_cast<cudaStream_t>(stream));
}

int csd_softmax_rows_f32_bf16(const float* logits, int in_pitch, void* probs, int out_pitch, int64_t rows, int cols,
                              float scale, csd_stream_t stream) {
  return csd::softmax_rows_launch<__nv_bfloat16>(logits, in_pitch, probs, out_pitch, rows, cols, scale,
                                                 static_cast<cudaStream_t>(stream));
}
int csd_softmax_rows_f32_f32(const float* logits, int in_pitch, void* probs, int out_pitch, int64_t rows, int cols,
                             float scale, csd_stream_t stream) {
  return csd::softmax_rows_launch<float>(logits, in_pitch, probs, out_pitch, rows, cols, scale,
                                         static_cast<cudaStream_t>(stream));
}

int csd_time_embedding_f32(const float* labels, int batch, int nf, int embedding_type, const float* fourier_w,
                           const float* w0, const float* b0, const float* w1, const float* b1, float* act_temb,
                           csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(labels && w0 && b0 && w1 && b1 && act_temb && batch >= 1 && nf >= 4, "time_embedding: bad arguments");
  CSD_REQUIRE(embedding_type == 0 || (embedding_type == 1 && fourier_w != nullptr), "time_embedding: bad embedding type");
  const size_t smem = sizeof(float) * ((embedding_type == 1 ? 2 * nf : nf) + 4 * nf);
  time_embedding_kernel<<<batch, 512, smem, static_cast<cudaStream_t>(stream)>>>(labels, nf, embedding_type, fourier_w, w0,
                                                                               b0, w1, b1, act_temb);
  CSD_LAUNCH_CHECK("time_embedding_kernel");
  return CSD_OK;
}

int csd_dense_rows_f32(const float* act_temb, const float* w, const float* bias, float* out, int batch, int in_dim,
                       int total_out, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(act_temb && w && out && batch >= 1 && in_dim >= 1 && total_out >= 1, "dense_rows: bad arguments");
  const dim3 grid((unsigned)ceil_div(total_out, kDenseTile), (unsigned)ceil_div(batch, kDenseTile));
  dense_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(act_temb, w, bias, out, batch, in_dim,
                                                                         total_out);
  CSD_LAUNCH_CHECK("dense_rows_kernel");
  return CSD_OK;
}

}  // extern "C"
