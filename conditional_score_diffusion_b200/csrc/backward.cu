// Backward-pass kernels of the score network that are not GEMM-shaped: GroupNorm(+SiLU) backward, FIR
// resampling adjoints, softmax backward, layout helpers (transpose, zero-stuffing, gradient layout), bias /
// time-embedding-projection gradients, the small fp32 GEMMs of the time-embedding MLP and the loss gradient.
//
// The reference gets all of these from PyTorch autograd (losses.py:345-407 `loss.backward()` through
// nn.GroupNorm / nn.SiLU / F.softmax / upfirdn2d's custom autograd.Function op/upfirdn2d.py:19-142). Each
// kernel below is the analytic adjoint of one forward kernel of this library; activation gradients are NHWC
// bf16 like the activations, parameter gradients fp32 in the reference's parameter layout.
#include "common.cuh"
#include "../../include/csd_b200.h"

#include <algorithm>

namespace csd {

static inline int flat_blocks(long long n, int per = 8) {
  long long b = ceil_div_ll(n, 256);
  long long cap = (long long)num_sms() * per;
  return (int)std::max<long long>(1, std::min(b, cap));
}

__device__ __forceinline__ float silu_grad(float u) {
  const float s = 1.f / (1.f + __expf(-u));
  return s * fmaf(u, 1.f - s, 1.f);
}

// ---- GroupNorm (+SiLU) backward -------------------------------------------------------------------------
// Forward: u = x * sc[b,c] + sh[b,c] (sc = rstd_g * gamma_c, sh = beta_c - mean_g * sc), y = silu(u) or u.
// Pass 1 (per source tensor): S[b, c] = (sum_p du, sum_p du * x) with du = dy * silu'(u).
struct GnBwdStatsParams {
  const bf16x8* x; int xv, xpv;            // source: vectors used, pitch in vectors
  const bf16x8* dy; int dy_pv, dy_voff;    // gradient of the GroupNorm output (concatenated channels)
  const float2* coef;                      // forward (scale, shift) of this source [batch, c, 2]
  float* s;                                // [batch, c_total, 2], this source starts at channel s_coff
  int c_total, s_coff;
  int hw, slabs, silu;
};

// Deterministic (no atomics): grid = (V / Vs, batch), a CTA owns Vs channel vectors of one image, walks all its pixels
// with ppb = blockDim / Vs pixel lanes, reduces the lanes in a fixed order and STORES the sums.
__global__ void __launch_bounds__(256) gn_bwd_stats_kernel(GnBwdStatsParams p, int Vs) {
  extern __shared__ float sm[];            // [ppb][16 * Vs]
  const int V = p.xv, C = V * 8;
  const int b = blockIdx.y;
  const int ppb = blockDim.x / Vs;
  const int v = threadIdx.x % Vs, pp = threadIdx.x / Vs;
  const int gv = blockIdx.x * Vs + v;
  const int n2 = 16 * Vs;
  float s1[8], s2[8], sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  if (pp < ppb && gv < V) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 c = p.coef[(long long)b * C + gv * 8 + i];
      sc[i] = c.x; sh[i] = c.y;
    }
    for (long long pix = pp; pix < p.hw; pix += ppb) {
      float fx[8], fd[8];
      unpack8(p.x[((long long)b * p.hw + pix) * p.xpv + gv], fx);
      unpack8(p.dy[((long long)b * p.hw + pix) * p.dy_pv + p.dy_voff + gv], fd);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float du = p.silu ? fd[i] * silu_grad(fmaf(fx[i], sc[i], sh[i])) : fd[i];
        s1[i] += du;
        s2[i] = fmaf(du, fx[i], s2[i]);
      }
    }
  }
  if (pp < ppb) {
    float* row = sm + (size_t)pp * n2 + v * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) { row[2 * i] = s1[i]; row[2 * i + 1] = s2[i]; }
  }
  __syncthreads();
  for (int c2 = threadIdx.x; c2 < n2; c2 += blockDim.x) {
    const int ch2 = blockIdx.x * n2 + c2;          // (channel, sum | sum*x) slot of this source
    if (ch2 >= 2 * C) continue;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int r = 0;
    for (; r + 3 < ppb; r += 4) {
      a0 += sm[(size_t)r * n2 + c2];
      a1 += sm[(size_t)(r + 1) * n2 + c2];
      a2 += sm[(size_t)(r + 2) * n2 + c2];
      a3 += sm[(size_t)(r + 3) * n2 + c2];
    }
    for (; r < ppb; ++r) a0 += sm[(size_t)r * n2 + c2];
    p.s[((long long)b * p.c_total + p.s_coff) * 2 + ch2] = (a0 + a1) + (a2 + a3);
  }
}

// Pass 2 (tiny): per (image, channel) coefficients of dx = A * du + B * x + Cc, from the forward sums (mean, rstd per
// group) and S. grid = batch. When `contrib` is set, S[b, c] is then overwritten with this image's contribution to
// (dgamma_c, dbeta_c) = (rstd * (S2 - mean * S1), S1), which gn_bwd_param_reduce_kernel sums over the batch.
__global__ void __launch_bounds__(256)
gn_bwd_coeffs_kernel(const float* __restrict__ sums0, int c0, const float* __restrict__ sums1, int c1,
                     const float* __restrict__ gamma, float* __restrict__ s, float4* __restrict__ coef, int hw,
                     int cpg, float eps, int contrib) {
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
    const float rstd = rsqrtf(var + eps);
    float m1 = 0.f, m2 = 0.f;
    for (int i = 0; i < cpg; ++i) {
      const int cc = g0 + i;
      const float s1 = s[((long long)b * C + cc) * 2], s2 = s[((long long)b * C + cc) * 2 + 1];
      m1 = fmaf(gamma[cc], s1, m1);
      m2 = fmaf(gamma[cc] * rstd, s2 - mean * s1, m2);
    }
    m1 *= inv_n;
    m2 *= inv_n;
    coef[(long long)b * C + c] = make_float4(rstd * gamma[c], -rstd * rstd * m2, -rstd * m1 + rstd * rstd * m2 * mean, rstd);
  }
  if (!contrib) return;
  __syncthreads();   // every thread has finished reading S (coef[..].w carries rstd; mean is recomputed)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g0 = (c / cpg) * cpg;
    float su = 0.f;
    for (int i = 0; i < cpg; ++i) {
      const int cc = g0 + i;
      su += (cc < c0) ? sums0[((long long)b * c0 + cc) * 2] : sums1[((long long)b * c1 + (cc - c0)) * 2];
    }
    const float mean = su * inv_n;
    const float rstd = coef[(long long)b * C + c].w;
    float* sp = s + ((long long)b * C + c) * 2;
    const float s1 = sp[0], s2 = sp[1];
    sp[0] = rstd * (s2 - mean * s1);
    sp[1] = s1;
  }
}

// dgamma[c] += sum_b contrib[b, c, 0], dbeta[c] += sum_b contrib[b, c, 1]. block (32, 8): 32 channels x 8 batch lanes,
// fixed-order tree (deterministic).
__global__ void __launch_bounds__(256)
gn_bwd_param_reduce_kernel(const float* __restrict__ contrib, float* __restrict__ dgamma, float* __restrict__ dbeta,
                           int batch, int C) {
  __shared__ float red[2][8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    for (int b = threadIdx.y; b < batch; b += 8) {
      const float2 v = *reinterpret_cast<const float2*>(contrib + ((long long)b * C + c) * 2);
      a0 += v.x;
      a1 += v.y;
    }
  }
  red[0][threadIdx.y][threadIdx.x] = a0;
  red[1][threadIdx.y][threadIdx.x] = a1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float g = 0.f, bt = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { g += red[0][k][threadIdx.x]; bt += red[1][k][threadIdx.x]; }
    dgamma[c] += g;
    dbeta[c] += bt;
  }
}

// Pass 3 (per source tensor): dx = A * dy * silu'(u) + B * x + C, written or accumulated into dx.
// grid = batch * slabs, block = V * ppb threads (V = 8-channel vectors per pixel): every thread keeps the coefficients
// of its 8 channels in registers and streams pixels (same decomposition as gn_apply_kernel).
struct GnBwdApplyParams {
  const bf16x8* x; int xv, xpv;
  const bf16x8* dy; int dy_pv, dy_voff;
  const float2* coef;        // forward (scale, shift) of this source [batch, c, 2]
  const float4* bcoef;       // backward (A, B, C, rstd) [batch, c_total, 4], this source starts at channel b_coff
  int c_total, b_coff;
  bf16x8* dx; int dx_pv;
  int hw, silu, accumulate, slabs;
};

__global__ void gn_bwd_apply_kernel(GnBwdApplyParams p) {
  const int V = p.xv, C = V * 8;
  const int b = blockIdx.x / p.slabs, slab = blockIdx.x % p.slabs;
  const int ppb = blockDim.x / V;
  const int v = threadIdx.x % V, pp = threadIdx.x / V;
  if (pp >= ppb) return;
  const long long chunk = ceil_div_ll(p.hw, p.slabs);
  const long long lo = slab * chunk, hi = min((long long)p.hw, lo + chunk);
  float sc[8], sh[8], ca[8], cb[8], cc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 fc = p.coef[(long long)b * C + v * 8 + i];
    const float4 bc = p.bcoef[(long long)b * p.c_total + p.b_coff + v * 8 + i];
    sc[i] = fc.x; sh[i] = fc.y; ca[i] = bc.x; cb[i] = bc.y; cc[i] = bc.z;
  }
  constexpr int U = 4;   // independent 16-byte loads in flight per tensor (the kernel is latency-bound otherwise)
  long long pix = lo + pp;
  for (; pix + (long long)(U - 1) * ppb < hi; pix += (long long)U * ppb) {
    bf16x8 vx[U], vd[U], vo[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long bp = (long long)b * p.hw + pix + (long long)u * ppb;
      vx[u] = p.x[bp * p.xpv + v];
      vd[u] = p.dy[bp * p.dy_pv + p.dy_voff + v];
      if (p.accumulate) vo[u] = p.dx[bp * p.dx_pv + v];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long bp = (long long)b * p.hw + pix + (long long)u * ppb;
      float fx[8], fd[8], o[8];
      unpack8(vx[u], fx);
      unpack8(vd[u], fd);
      if (p.accumulate) unpack8(vo[u], o);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float du = p.silu ? fd[i] * silu_grad(fmaf(fx[i], sc[i], sh[i])) : fd[i];
        o[i] += fmaf(ca[i], du, fmaf(cb[i], fx[i], cc[i]));
      }
      p.dx[bp * p.dx_pv + v] = pack8(o);
    }
  }
  for (; pix < hi; pix += ppb) {
    const long long bp = (long long)b * p.hw + pix;
    float fx[8], fd[8], o[8];
    unpack8(p.x[bp * p.xpv + v], fx);
    unpack8(p.dy[bp * p.dy_pv + p.dy_voff + v], fd);
    if (p.accumulate) unpack8(p.dx[bp * p.dx_pv + v], o);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float du = p.silu ? fd[i] * silu_grad(fmaf(fx[i], sc[i], sh[i])) : fd[i];
      o[i] += fmaf(ca[i], du, fmaf(cb[i], fx[i], cc[i]));
    }
    p.dx[bp * p.dx_pv + v] = pack8(o);
  }
}

// ---- FIR resampling adjoints ------------------------------------------------------------------------------
// mode 1 (adjoint of up x2, pad (2,1)):   din[i] = sum_k kf[k] * g[2i - 1 + k]          (g: 2h x 2w -> h x w)
// mode 2 (adjoint of down x2, pad (1,1)): din[2i] = g[i-1] kf[0] + g[i] kf[2]; din[2i+1] = g[i] kf[1] + g[i+1] kf[3]
// mode 3 (adjoint of the pad-(2,2) pre-filter): din[j] = sum_k kf[k] * g[j - 1 + k]     (g: (h+1) x (w+1) -> h x w)
// kf holds the per-axis taps of the ADJOINT (already scaled). One thread = one input-gradient pixel x 8 channels.
struct FirBwdParams {
  const bf16x8* g; bf16x8* din;
  int batch, h, w, gh, gw, cvec;   // (h, w): forward input extent; (gh, gw): forward output extent
  float kf[4];
  int mode, accumulate;
};

__global__ void __launch_bounds__(256) fir_bwd_kernel(FirBwdParams p) {
  const long long total = (long long)p.batch * p.h * p.w * p.cvec;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % p.cvec);
    long long pix = idx / p.cvec;
    const int x = (int)(pix % p.w);
    const int y = (int)((pix / p.w) % p.h);
    const int b = (int)(pix / ((long long)p.w * p.h));
    int iy[4], ix[4], ny, nx;
    float wy[4], wx[4];
    if (p.mode == 1) {
      ny = nx = 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) { iy[k] = 2 * y - 1 + k; wy[k] = p.kf[k]; ix[k] = 2 * x - 1 + k; wx[k] = p.kf[k]; }
    } else if (p.mode == 2) {
      ny = nx = 2;
      const int i = y >> 1, j = x >> 1;
      if ((y & 1) == 0) { iy[0] = i - 1; wy[0] = p.kf[0]; iy[1] = i; wy[1] = p.kf[2]; }
      else              { iy[0] = i;     wy[0] = p.kf[1]; iy[1] = i + 1; wy[1] = p.kf[3]; }
      if ((x & 1) == 0) { ix[0] = j - 1; wx[0] = p.kf[0]; ix[1] = j; wx[1] = p.kf[2]; }
      else              { ix[0] = j;     wx[0] = p.kf[1]; ix[1] = j + 1; wx[1] = p.kf[3]; }
    } else {
      ny = nx = 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) { iy[k] = y - 1 + k; wy[k] = p.kf[k]; ix[k] = x - 1 + k; wx[k] = p.kf[k]; }
    }
    float acc[8];
    if (p.accumulate) unpack8(p.din[idx], acc);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (a >= ny) break;
      if (iy[a] < 0 || iy[a] >= p.gh) continue;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c >= nx) break;
        if (ix[c] < 0 || ix[c] >= p.gw) continue;
        float f[8];
        unpack8(p.g[(((long long)b * p.gh + iy[a]) * p.gw + ix[c]) * p.cvec + cv], f);
        const float wgt = wy[a] * wx[c];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(wgt, f[i], acc[i]);
      }
    }
    p.din[idx] = pack8(acc);
  }
}

// ---- softmax backward: ds = scale * p * (dp - sum_j p_j dp_j), one warp per row ---------------------------------
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const __nv_bfloat16* __restrict__ probs, int p_pitch, const float* __restrict__ dp, int dp_pitch,
                   __nv_bfloat16* __restrict__ ds, int ds_pitch, long long rows, int cols, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __nv_bfloat16* pr = probs + row * p_pitch;
  const float* dr = dp + row * dp_pitch;
  float dot = 0.f;
  for (int c = lane; c < cols; c += 32) dot = fmaf(__bfloat162float(pr[c]), dr[c], dot);
  dot = warp_sum(dot);
  __nv_bfloat16* out = ds + row * ds_pitch;
  for (int c = lane; c < ds_pitch; c += 32)
    out[c] = __float2bfloat16_rn(c < cols ? scale * __bfloat162float(pr[c]) * (dr[c] - dot) : 0.f);
}

// ---- batched bf16 transpose: out[z, c, r] = in[z, r, c] ------------------------------------------------------
__global__ void __launch_bounds__(256)
transpose_kernel(const __nv_bfloat16* __restrict__ in, int in_pitch, long long in_z, __nv_bfloat16* __restrict__ out,
                 int out_pitch, long long out_z, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int z = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + tx;
    tile[k][tx] = (r < rows && c < cols) ? in[z * in_z + (long long)r * in_pitch + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + tx;
    if (c < cols && r < out_pitch)   // columns in [rows, out_pitch) are the zero padding of the K-major operand
      out[z * out_z + (long long)c * out_pitch + r] = (r < rows) ? tile[tx][k] : __float2bfloat16_rn(0.f);
  }
}

// ---- dst = alpha * src (+ dst) over bf16 vectors ---------------------------------------------------------------
__global__ void __launch_bounds__(256)
axpy_bf16_kernel(const bf16x8* __restrict__ src, bf16x8* dst, long long nvec, float alpha, int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float a[8], d[8];
    unpack8(src[i], a);
    if (accumulate) {
      unpack8(dst[i], d);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = fmaf(alpha, a[k], d[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = alpha * a[k];
    }
    dst[i] = pack8(d);
  }
}

// ---- dropout: out = keep(seed, salt, i) ? x / (1 - p) : 0, mask recomputed from a counter-based hash ------------
// The same call applied to the gradient is the backward pass (no mask tensor is stored). nn.Dropout in
// ResnetBlockBigGANpp / ResnetBlockDDPM (models/layerspp.py:266, models/layers.py:664) draws from torch's
// generator; here the per-step seed is drawn from it too (a device int64), the per-element stream is our own.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
dropout_bf16_kernel(const bf16x8* __restrict__ x, bf16x8* __restrict__ out, long long nvec, float p,
                    const long long* __restrict__ seed, unsigned long long salt) {
  const unsigned long long s = mix64((unsigned long long)(*seed) ^ (salt * 0x9E3779B97F4A7C15ULL));
  const unsigned int thresh = (unsigned int)(p * 65536.f);
  const float inv_keep = 1.f / (1.f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long r0 = mix64(s + 2ULL * (unsigned long long)i);
    const unsigned long long r1 = mix64(s + 2ULL * (unsigned long long)i + 1ULL);
    float f[8];
    unpack8(x[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned int u = (unsigned int)(((k < 4 ? r0 : r1) >> (16 * (k & 3))) & 0xFFFFu);
      f[k] = u >= thresh ? f[k] * inv_keep : 0.f;
    }
    out[i] = pack8(f);
  }
}

// ---- zero stuffing: dst[b, y*s+off, x*s+off, :] = src[b, y, x, :], zero elsewhere --------------------------------
__global__ void __launch_bounds__(256)
zero_stuff_kernel(const bf16x8* __restrict__ src, bf16x8* __restrict__ dst, int batch, int h, int w, int dh, int dw,
                  int cvec, int stride, int offset) {
  const long long total = (long long)batch * dh * dw * cvec;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvec);
    long long pix = idx / cvec;
    const int x = (int)(pix % dw);
    const int y = (int)((pix / dw) % dh);
    const int b = (int)(pix / ((long long)dw * dh));
    const int sy = y - offset, sx = x - offset;
    bf16x8 v;
    *reinterpret_cast<uint4*>(&v) = make_uint4(0, 0, 0, 0);
    if (sy >= 0 && sx >= 0 && sy % stride == 0 && sx % stride == 0 && sy / stride < h && sx / stride < w)
      v = src[(((long long)b * h + sy / stride) * w + sx / stride) * cvec + cv];
    dst[idx] = v;
  }
}

// ---- gradient layout: NCHW fp32 (one or two tensors, per-sample scale) -> NHWC bf16 ----------------------------
__global__ void __launch_bounds__(256)
nchw_grad_to_nhwc_kernel(const float* __restrict__ s0, int c0, const float* __restrict__ rs0, const float* __restrict__ s1,
                         int c1, const float* __restrict__ rs1, bf16x8* __restrict__ out, int cvec, int batch, long long hw) {
  const long long total = (long long)batch * hw * cvec;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long pix = idx % hw;
    const int cv = (int)((idx / hw) % cvec);
    const int b = (int)(idx / (hw * cvec));
    const float r0 = rs0 != nullptr ? __ldg(rs0 + b) : 1.f;
    const float r1 = rs1 != nullptr ? __ldg(rs1 + b) : 1.f;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = cv * 8 + i;
      float v = 0.f;
      if (c < c0) { if (s0 != nullptr) v = __ldg(s0 + ((long long)b * c0 + c) * hw + pix) * r0; }
      else if (c < c0 + c1) { if (s1 != nullptr) v = __ldg(s1 + ((long long)b * c1 + (c - c0)) * hw + pix) * r1; }
      f[i] = v;
    }
    out[((long long)b * hw + pix) * cvec + cv] = pack8(f);
  }
}

// ---- bias / time-embedding projection gradients from per-(image, channel) sums of the output gradient -----------
// Block = 32 channels x 8 batch lanes: the per-image loads (and the dtproj read-modify-writes, one owner per (b, ch))
// of the lanes are independent, so the batch loop is ~batch/8 deep instead of a serial chain of `batch` L2 round trips
// (14 us per launch x 73 launches per training step as a one-thread-per-channel loop, ncu). The per-channel bias sum
// is combined over the lanes in a fixed order.
__global__ void __launch_bounds__(256)
bias_temb_grad_kernel(const float* __restrict__ sums, int sums_c, int batch, int c, float scale, float* dbias0,
                      float* dbias1, float* dtproj, int tpitch) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, by = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (ch < c) {
    for (int b = by; b < batch; b += 8) {
      const float v = sums[((long long)b * sums_c + ch) * 2] * scale;
      acc += v;
      if (dtproj != nullptr) dtproj[(long long)b * tpitch + ch] += v;
    }
  }
  red[by][lane] = acc;
  __syncthreads();
  if (by == 0 && ch < c) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += red[k][lane];
    if (dbias0 != nullptr) dbias0[ch] += a;
    if (dbias1 != nullptr) dbias1[ch] += a;
  }
}

// ---- small fp32 GEMM (time-embedding MLP and Dense_0 projections; <= 1 GFLOP per call) ---------------------------
// C[m, n] = alpha * sum_k opA(m, k) * opB(k, n) + beta * C[m, n] + bias[n];  opA(m,k) = ta ? A[k*lda+m] : A[m*lda+k],
// opB(k,n) = tb ? B[n*ldb+k] : B[k*ldb+n].
__global__ void __launch_bounds__(256)
sgemm_small_kernel(int ta, int tb, int m, int n, int k, float alpha, const float* __restrict__ a, int lda,
                   const float* __restrict__ b, int ldb, float beta, float* c, int ldc, const float* __restrict__ bias) {
  const long long total = (long long)m * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx % n), row = (int)(idx / n);
    float acc = 0.f;
    for (int kk = 0; kk < k; ++kk) {
      const float av = ta ? a[(long long)kk * lda + row] : a[(long long)row * lda + kk];
      const float bv = tb ? b[(long long)col * ldb + kk] : b[(long long)kk * ldb + col];
      acc = fmaf(av, bv, acc);
    }
    float r = alpha * acc;
    if (beta != 0.f) r = fmaf(beta, c[(long long)row * ldc + col], r);
    if (bias != nullptr) r += bias[col];
    c[(long long)row * ldc + col] = r;
  }
}

// Long-K variant: one warp per output element, lanes stride over k (the d act_temb = dtproj @ W contraction has
// K = sum of all Dense_0 widths ~ 5-16 k and only batch x 4nf outputs: a thread-per-output loop would be one long
// dependent chain).
__global__ void __launch_bounds__(256)
sgemm_longk_kernel(int ta, int tb, int m, int n, int k, float alpha, const float* __restrict__ a, int lda,
                   const float* __restrict__ b, int ldb, float beta, float* c, int ldc, const float* __restrict__ bias) {
  const int lane = threadIdx.x & 31;
  const long long total = (long long)m * n;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long idx = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; idx < total; idx += warps) {
    const int col = (int)(idx % n), row = (int)(idx / n);
    float acc = 0.f;
    for (int kk = lane; kk < k; kk += 32) {
      const float av = ta ? a[(long long)kk * lda + row] : a[(long long)row * lda + kk];
      const float bv = tb ? b[(long long)col * ldb + kk] : b[(long long)kk * ldb + col];
      acc = fmaf(av, bv, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float r = alpha * acc;
      if (beta != 0.f) r = fmaf(beta, c[(long long)row * ldc + col], r);
      if (bias != nullptr) r += bias[col];
      c[(long long)row * ldc + col] = r;
    }
  }
}

// y = silu(x) (grad == 0) or y = dy * silu'(x) (grad == 1), full-precision exp like the forward time-embedding kernel
__global__ void __launch_bounds__(256)
silu_f32_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ y, long long n, int grad) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float s = 1.f / (1.f + expf(-v));
    y[i] = grad ? dy[i] * s * fmaf(v, 1.f - s, 1.f) : v * s;
  }
}

// emb[b, :] of models/layers.py:524-538 (positional) / models/layerspp.py:32-41 (Gaussian Fourier)
__global__ void __launch_bounds__(128)
time_features_kernel(const float* __restrict__ labels, int nf, int embedding_type, const float* __restrict__ fourier_w,
                     float* __restrict__ emb) {
  const int b = blockIdx.x;
  const float t = labels[b];
  if (embedding_type == 1) {
    float* e = emb + (long long)b * 2 * nf;
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
      const float a = t * fourier_w[i] * 2.f * 3.14159265358979323846f;
      e[i] = sinf(a);
      e[nf + i] = cosf(a);
    }
  } else {
    float* e = emb + (long long)b * nf;
    const int half = nf / 2;
    const float c = logf(10000.f) / (float)(half - 1);
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
      const float a = t * expf(-c * (float)i);
      e[i] = sinf(a);
      e[half + i] = cosf(a);
    }
    if ((nf & 1) && threadIdx.x == 0) e[nf - 1] = 0.f;
  }
}

// dscore[b, i] = gl[b] * 2 * w[b] * a[b] * (a[b] * score[b, i] + c[b] * z[b, i])   (adjoint of csd_dsm_loss_f32)
__global__ void __launch_bounds__(256)
dsm_loss_bwd_kernel(const float* __restrict__ score, const float* __restrict__ z, const float* __restrict__ a,
                    const float* __restrict__ c, const float* __restrict__ w, const float* __restrict__ gl,
                    float* __restrict__ dscore, long long per_sample) {
  const int b = blockIdx.y;
  const float ab = a[b], cb = c[b];
  const float k = 2.f * w[b] * ab * gl[b];
  const long long base = (long long)b * per_sample;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x)
    dscore[base + i] = k * fmaf(ab, score[base + i], cb * z[base + i]);
}

}  // namespace csd

extern "C" {

int csd_gn_bwd_stats_bf16(const void* x, int c, int x_pitch, const void* dy, int dy_pitch, int dy_c_off,
                          const float* fwd_coef, float* s, int c_total, int s_c_off, int batch, int hw, int silu,
                          csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && dy && fwd_coef && s, "gn_bwd_stats: null pointer");
  CSD_REQUIRE(c >= 8 && c % 8 == 0 && x_pitch % 8 == 0 && dy_pitch % 8 == 0 && dy_c_off % 8 == 0 && c <= 8192,
              "gn_bwd_stats: channels / pitches must be multiples of 8 (c=%d)", c);
  GnBwdStatsParams p;
  p.x = static_cast<const bf16x8*>(x); p.xv = c / 8; p.xpv = x_pitch / 8;
  p.dy = static_cast<const bf16x8*>(dy); p.dy_pv = dy_pitch / 8; p.dy_voff = dy_c_off / 8;
  p.coef = reinterpret_cast<const float2*>(fwd_coef);
  p.s = s; p.c_total = c_total; p.s_coff = s_c_off; p.hw = hw; p.silu = silu;
  const int V = c / 8;
  p.slabs = 1;
  // Vs channel vectors per CTA: 32-byte pixel rows at least, fatter CTAs once the batch alone fills the machine
  int vs = 2;
  while (vs * 2 <= 8 && V % (vs * 2) == 0 && (long long)batch * (V / (vs * 2)) >= 4LL * num_sms()) vs *= 2;
  if (V % vs != 0) vs = 1;
  const int ppb = 256 / vs;
  const size_t smem = sizeof(float) * 16 * vs * (size_t)ppb;
  CSD_REQUIRE(batch <= 65535, "gn_bwd_stats: batch %d too large", batch);
  gn_bwd_stats_kernel<<<dim3((unsigned)(V / vs), (unsigned)batch), vs * ppb, smem, static_cast<cudaStream_t>(stream)>>>(p, vs);
  CSD_LAUNCH_CHECK("gn_bwd_stats_kernel");
  return CSD_OK;
}

int csd_gn_bwd_coeffs_f32(const float* sums0, int c0, const float* sums1, int c1, const float* gamma, float* s,
                          float* bwd_coef, float* dgamma, float* dbeta, int batch, int hw, int groups, float eps,
                          csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(sums0 && gamma && s && bwd_coef, "gn_bwd_coeffs: null pointer");
  if (sums1 == nullptr) c1 = 0;
  const int C = c0 + c1;
  CSD_REQUIRE(groups >= 1 && C % groups == 0, "gn_bwd_coeffs: %d channels not divisible by %d groups", C, groups);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int contrib = (dgamma != nullptr && dbeta != nullptr) ? 1 : 0;
  gn_bwd_coeffs_kernel<<<batch, 256, 0, st>>>(sums0, c0, sums1, c1, gamma, s, reinterpret_cast<float4*>(bwd_coef), hw,
                                              C / groups, eps, contrib);
  CSD_LAUNCH_CHECK("gn_bwd_coeffs_kernel");
  if (contrib) {
    gn_bwd_param_reduce_kernel<<<ceil_div(C, 32), dim3(32, 8), 0, st>>>(s, dgamma, dbeta, batch, C);
    CSD_LAUNCH_CHECK("gn_bwd_param_reduce_kernel");
  }
  return CSD_OK;
}

int csd_gn_bwd_apply_bf16(const void* x, int c, int x_pitch, const void* dy, int dy_pitch, int dy_c_off,
                          const float* fwd_coef, const float* bwd_coef, int c_total, int b_c_off, void* dx, int dx_pitch,
                          int batch, int hw, int silu, int accumulate, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && dy && fwd_coef && bwd_coef && dx, "gn_bwd_apply: null pointer");
  CSD_REQUIRE(c >= 8 && c % 8 == 0 && x_pitch % 8 == 0 && dy_pitch % 8 == 0 && dy_c_off % 8 == 0 && dx_pitch % 8 == 0,
              "gn_bwd_apply: channels / pitches must be multiples of 8 (c=%d)", c);
  GnBwdApplyParams p;
  p.x = static_cast<const bf16x8*>(x); p.xv = c / 8; p.xpv = x_pitch / 8;
  p.dy = static_cast<const bf16x8*>(dy); p.dy_pv = dy_pitch / 8; p.dy_voff = dy_c_off / 8;
  p.coef = reinterpret_cast<const float2*>(fwd_coef);
  p.bcoef = reinterpret_cast<const float4*>(bwd_coef);
  p.c_total = c_total; p.b_coff = b_c_off;
  p.dx = static_cast<bf16x8*>(dx); p.dx_pv = dx_pitch / 8;
  p.hw = hw; p.silu = silu; p.accumulate = accumulate;
  const int V = c / 8;
  CSD_REQUIRE(V <= 1024, "gn_bwd_apply: %d channels exceed the 8192 supported", c);
  const int ppb = std::max(1, 256 / V);
  int slabs = ceil_div(num_sms() * 16, batch);
  p.slabs = std::max(1, std::min(slabs, std::max(1, hw / (ppb * 4))));
  gn_bwd_apply_kernel<<<batch * p.slabs, V * ppb, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CSD_LAUNCH_CHECK("gn_bwd_apply_kernel");
  return CSD_OK;
}

int csd_fir_resample_bwd_nhwc_bf16(const void* g, void* din, int batch, int h, int w, int c_pitch, int mode,
                                   const float* taps4_host, int accumulate, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(g && din && taps4_host, "fir_resample_bwd: null pointer");
  CSD_REQUIRE(mode >= 1 && mode <= 3, "fir_resample_bwd: mode %d", mode);
  CSD_REQUIRE(c_pitch % 8 == 0, "fir_resample_bwd: channel pitch %d not a multiple of 8", c_pitch);
  FirBwdParams p;
  p.g = static_cast<const bf16x8*>(g);
  p.din = static_cast<bf16x8*>(din);
  p.batch = batch; p.h = h; p.w = w; p.cvec = c_pitch / 8;
  p.gh = mode == 1 ? 2 * h : (mode == 2 ? h / 2 : h + 1);
  p.gw = mode == 1 ? 2 * w : (mode == 2 ? w / 2 : w + 1);
  p.mode = mode; p.accumulate = accumulate;
  float sum = 0.f;
  for (int i = 0; i < 4; ++i) sum += taps4_host[i];
  CSD_REQUIRE(sum != 0.f, "fir_resample_bwd: taps sum to zero");
  // the forward kernels use the flipped normalised taps f[i] = taps[3-i]/sum (x2 per axis when upsampling); the
  // adjoint correlates with the un-flipped filter
  for (int i = 0; i < 4; ++i) p.kf[i] = taps4_host[i] / sum * (mode == 1 ? 2.f : 1.f);
  const long long total = (long long)batch * h * w * p.cvec;
  if (total == 0) return CSD_OK;
  fir_bwd_kernel<<<flat_blocks(total, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CSD_LAUNCH_CHECK("fir_bwd_kernel");
  return CSD_OK;
}

int csd_softmax_bwd_bf16(const void* probs, int p_pitch, const float* dp, int dp_pitch, void* ds, int ds_pitch,
                         int64_t rows, int cols, float scale, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(probs && dp && ds && rows >= 1 && cols >= 1, "softmax_bwd: bad arguments");
  softmax_bwd_kernel<<<(unsigned)ceil_div_ll(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(probs), p_pitch, dp, dp_pitch, static_cast<__nv_bfloat16*>(ds), ds_pitch, rows,
      cols, scale);
  CSD_LAUNCH_CHECK("softmax_bwd_kernel");
  return CSD_OK;
}

int csd_transpose_bf16(const void* in, int in_pitch, int64_t in_z_stride, void* out, int out_pitch, int64_t out_z_stride,
                       int rows, int cols, int z, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(in && out && rows >= 1 && cols >= 1 && z >= 1 && z <= 65535, "transpose: bad arguments");
  CSD_REQUIRE(out_pitch >= rows && in_pitch >= cols, "transpose: pitches smaller than the matrix");
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32), (unsigned)z);
  transpose_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(in), in_pitch,
                                                                      in_z_stride, static_cast<__nv_bfloat16*>(out),
                                                                      out_pitch, out_z_stride, rows, cols);
  CSD_LAUNCH_CHECK("transpose_kernel");
  return CSD_OK;
}

int csd_axpy_bf16(const void* src, void* dst, int64_t n, float alpha, int accumulate, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(src && dst && n >= 0 && n % 8 == 0, "axpy_bf16: element count must be a multiple of 8");
  if (n == 0) return CSD_OK;
  axpy_bf16_kernel<<<flat_blocks(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16x8*>(src), static_cast<bf16x8*>(dst), n / 8, alpha, accumulate);
  CSD_LAUNCH_CHECK("axpy_bf16_kernel");
  return CSD_OK;
}

int csd_dropout_bf16(const void* x, void* out, int64_t n, float p, const int64_t* seed_dev, uint64_t salt,
                     csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && out && seed_dev && n >= 0 && n % 8 == 0, "dropout: element count must be a multiple of 8");
  CSD_REQUIRE(p >= 0.f && p < 1.f, "dropout: p = %f out of [0, 1)", (double)p);
  if (n == 0) return CSD_OK;
  dropout_bf16_kernel<<<flat_blocks(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16x8*>(x), static_cast<bf16x8*>(out), n / 8, p, reinterpret_cast<const long long*>(seed_dev),
      (unsigned long long)salt);
  CSD_LAUNCH_CHECK("dropout_bf16_kernel");
  return CSD_OK;
}

int csd_zero_stuff_nhwc_bf16(const void* src, void* dst, int batch, int h, int w, int dst_h, int dst_w, int c_pitch,
                             int stride, int offset, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(src && dst && c_pitch % 8 == 0 && stride >= 1 && offset >= 0, "zero_stuff: bad arguments");
  CSD_REQUIRE((h - 1) * stride + offset < dst_h && (w - 1) * stride + offset < dst_w, "zero_stuff: destination too small");
  const long long total = (long long)batch * dst_h * dst_w * (c_pitch / 8);
  zero_stuff_kernel<<<flat_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16x8*>(src), static_cast<bf16x8*>(dst), batch, h, w, dst_h, dst_w, c_pitch / 8, stride, offset);
  CSD_LAUNCH_CHECK("zero_stuff_kernel");
  return CSD_OK;
}

int csd_nchw_grad_to_nhwc_bf16(const float* g0, int c0, const float* row_scale0, const float* g1, int c1,
                               const float* row_scale1, void* out, int c_pad, int batch, int h, int w,
                               csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(out && c_pad % 8 == 0 && c0 + c1 <= c_pad && c0 >= 1 && c1 >= 0, "nchw_grad_to_nhwc: bad arguments");
  const long long total = (long long)batch * h * w * (c_pad / 8);
  nchw_grad_to_nhwc_kernel<<<flat_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      g0, c0, row_scale0, g1, c1, row_scale1, static_cast<bf16x8*>(out), c_pad / 8, batch, (long long)h * w);
  CSD_LAUNCH_CHECK("nchw_grad_to_nhwc_kernel");
  return CSD_OK;
}

int csd_bias_temb_grad_f32(const float* chan_sums, int sums_c, int batch, int c, float scale, float* dbias0,
                           float* dbias1, float* dtproj, int tproj_pitch, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(chan_sums && batch >= 1 && c >= 1 && sums_c >= c, "bias_temb_grad: bad arguments");
  bias_temb_grad_kernel<<<ceil_div(c, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(chan_sums, sums_c, batch, c, scale,
                                                                                     dbias0, dbias1, dtproj, tproj_pitch);
  CSD_LAUNCH_CHECK("bias_temb_grad_kernel");
  return CSD_OK;
}

int csd_sgemm_small_f32(int trans_a, int trans_b, int m, int n, int k, float alpha, const float* a, int lda,
                        const float* b, int ldb, float beta, float* c, int ldc, const float* bias, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(a && b && c && m >= 1 && n >= 1 && k >= 1, "sgemm_small: bad arguments");
  if (k >= 512 && (long long)m * n <= (1 << 17)) {
    sgemm_longk_kernel<<<flat_blocks((long long)m * n * 32, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        trans_a, trans_b, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, bias);
    CSD_LAUNCH_CHECK("sgemm_longk_kernel");
    return CSD_OK;
  }
  sgemm_small_kernel<<<flat_blocks((long long)m * n, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      trans_a, trans_b, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, bias);
  CSD_LAUNCH_CHECK("sgemm_small_kernel");
  return CSD_OK;
}

int csd_silu_f32(const float* x, const float* dy, float* y, int64_t n, int grad, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && y && n >= 1 && (grad == 0 || dy != nullptr), "silu_f32: bad arguments");
  silu_f32_kernel<<<flat_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, y, n, grad);
  CSD_LAUNCH_CHECK("silu_f32_kernel");
  return CSD_OK;
}

int csd_time_features_f32(const float* labels, int batch, int nf, int embedding_type, const float* fourier_w, float* emb,
                          csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(labels && emb && batch >= 1 && nf >= 2, "time_features: bad arguments");
  CSD_REQUIRE(embedding_type == 0 || (embedding_type == 1 && fourier_w != nullptr), "time_features: embedding type %d",
              embedding_type);
  time_features_kernel<<<batch, 128, 0, static_cast<cudaStream_t>(stream)>>>(labels, nf, embedding_type, fourier_w, emb);
  CSD_LAUNCH_CHECK("time_features_kernel");
  return CSD_OK;
}

int csd_dsm_loss_bwd_f32(const float* score, const float* z, const float* a, const float* c, const float* w,
                         const float* grad_losses, float* dscore, int batch, int64_t per_sample, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(score && z && a && c && w && grad_losses && dscore && batch >= 1 && batch <= 65535 && per_sample >= 1,
              "dsm_loss_bwd: bad arguments");
  const long long bx = std::max<long long>(1, std::min<long long>(ceil_div_ll(per_sample, 256),
                                                                   ceil_div_ll((long long)num_sms() * 8, batch)));
  dsm_loss_bwd_kernel<<<dim3((unsigned)bx, (unsigned)batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      score, z, a, c, w, grad_losses, dscore, per_sample);
  CSD_LAUNCH_CHECK("dsm_loss_bwd_kernel");
  return CSD_OK;
}

}  // extern "C"
