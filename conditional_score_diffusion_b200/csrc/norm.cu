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

namespace csd {

struct GnParams {
  const bf16x8* src0;
  const bf16x8* src1;
  int v0, v1;          // 8-channel vectors taken from each source
  int pv0, pv1;        // pitch of each source in vectors
  int batch, hw, groups, cpg;
  float* sums0;        // per-channel (sum, sum of squares): [batch, c0, 2]
  float* sums1;        //                                     [batch, c1, 2]
  const float* gamma;
  const float* beta;
  bf16x8* out;
  int out_pv;
  float eps;
  int silu;
  int slabs;
};

__device__ __forceinline__ bf16x8 gn_load(const GnParams& p, int b, long long pix, int v) {
  if (v < p.v0) return p.src0[((long long)b * p.hw + pix) * p.pv0 + v];
  return p.src1[((long long)b * p.hw + pix) * p.pv1 + (v - p.v0)];
}

// Per-channel statistics of ONE tensor. grid = batch * slabs; block = V * PPB threads (V = vectors per
// pixel). Dynamic smem: 2*C floats. Each tensor's sums are computed once and shared by every GroupNorm
// that reads the tensor (the down-path activations are normalised twice: by the next block and, through
// the skip concatenation, by the up path).
__global__ void gn_chan_stats_kernel(GnParams p) {
  extern __shared__ float sm[];     // [ppb][2*C] partial sums, reduced by a fixed-order tree (no shared atomics)
  const int V = p.v0, C = V * 8;
  const int b = blockIdx.x / p.slabs, slab = blockIdx.x % p.slabs;
  const int ppb = blockDim.x / V;
  const int v = threadIdx.x % V, pp = threadIdx.x / V;
  const long long chunk = ceil_div_ll(p.hw, p.slabs);
  const long long lo = slab * chunk, hi = min((long long)p.hw, lo + chunk);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (pp < ppb) {
    long long pix = lo + pp;
    for (; pix + 3LL * ppb < hi; pix += 4LL * ppb) {   // 4 independent 16-byte loads in flight
      bf16x8 v4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v4[u] = gn_load(p, b, pix + (long long)u * ppb, v);
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
      unpack8(gn_load(p, b, pix, v), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] = fmaf(f[i], f[i], q[i]);
      }
    }
    float* row = sm + (size_t)pp * 2 * C + v * 16;   // (sum, sumsq) interleaved per channel
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      row[2 * i] = s[i];
      row[2 * i + 1] = q[i];
    }
  }
  __syncthreads();
  for (int c2 = threadIdx.x; c2 < 2 * C; c2 += blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < ppb; ++r) a += sm[(size_t)r * 2 * C + c2];
    atomicAdd(p.sums0 + (long long)b * 2 * C + c2, a);
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
__global__ void gn_apply_kernel(GnParams p) {
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
    bf16x8 v4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v4[u] = gn_load(p, b, pix + (long long)u * ppb, v);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(v4[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = fmaf(f[i], sc[i], sh[i]);
        f[i] = p.silu ? silu_f(y) : y;
      }
      p.out[((long long)b * p.hw + pix + (long long)u * ppb) * p.out_pv + v] = pack8(f);
    }
  }
  for (; pix < hi; pix += ppb) {
    float f[8];
    unpack8(gn_load(p, b, pix, v), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = fmaf(f[i], sc[i], sh[i]);
      f[i] = p.silu ? silu_f(y) : y;
    }
    p.out[((long long)b * p.hw + pix) * p.out_pv + v] = pack8(f);
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

static int gn_fill(GnParams& p, const void* src0, int c0, int pitch0, const void* src1, int c1, int pitch1, int batch,
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
  p.src0 = static_cast<const bf16x8*>(src0);
  p.src1 = static_cast<const bf16x8*>(src1);
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

// ---- layout conversion ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ s0, int c0, const float* __restrict__ s1, int c1, bf16x8* __restrict__ out,
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
      f[i] = v;
    }
    out[((long long)b * hw + pix) * cvec + cv] = pack8(f);
  }
}

__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, int pitch, int c_off, int c_cnt, float* __restrict__ dst,
                    int batch, long long hw, const float* __restrict__ row_scale) {
  const long long total = (long long)batch * c_cnt * hw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long pix = idx % hw;
    const int c = (int)((idx / hw) % c_cnt);
    const int b = (int)(idx / (hw * c_cnt));
    float v = __bfloat162float(src[((long long)b * hw + pix) * pitch + c_off + c]);
    if (row_scale != nullptr) v *= __ldg(row_scale + b);
    dst[idx] = v;
  }
}

// ---- softmax ---------------------------------------------------------------------------------------
// One warp per row; rows are short (<= 1024 keys), so the three passes hit L1.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ logits, int in_pitch, __nv_bfloat16* __restrict__ probs, int out_pitch,
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
  __nv_bfloat16* out = probs + row * out_pitch;
  for (int c = lane; c < out_pitch; c += 32)
    out[c] = __float2bfloat16_rn(c < cols ? __expf(in[c] * scale - m) * inv : 0.f);
}

// ---- time embedding ----------------------------------------------------------------------------------
// One CTA per batch row. emb -> Linear -> SiLU -> Linear -> SiLU (the SiLU every block applies to temb
// before its Dense_0, models/layerspp.py:263, is hoisted here).
__global__ void __launch_bounds__(256)
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
  for (int j = threadIdx.x; j < hid; j += blockDim.x) {
    float acc = b0[j];
    const float* wr = w0 + (long long)j * embed;
    for (int k = 0; k < embed; ++k) acc = fmaf(wr[k], emb[k], acc);
    h0[j] = acc / (1.f + expf(-acc));  // SiLU, full-precision exp
  }
  __syncthreads();
  for (int j = threadIdx.x; j < hid; j += blockDim.x) {
    float acc = b1[j];
    const float* wr = w1 + (long long)j * hid;
    for (int k = 0; k < hid; ++k) acc = fmaf(wr[k], h0[k], acc);
    act_temb[(long long)b * hid + j] = acc / (1.f + expf(-acc));
  }
}

// One warp per output row j: the weight row stays in registers, the batch loop reuses it.
__global__ void __launch_bounds__(256)
dense_rows_kernel(const float* __restrict__ act, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, int batch, int in_dim, int total_out) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= total_out) return;
  constexpr int MAXR = 32;  // in_dim <= 1024
  float wr[MAXR];
  const int per = ceil_div(in_dim, 32);
#pragma unroll
  for (int i = 0; i < MAXR; ++i) {
    const int k = lane + 32 * i;
    wr[i] = (i < per && k < in_dim) ? __ldg(w + (long long)j * in_dim + k) : 0.f;
  }
  const float bj = bias != nullptr ? bias[j] : 0.f;
  for (int b = 0; b < batch; ++b) {
    const float* a = act + (long long)b * in_dim;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < MAXR; ++i) {
      const int k = lane + 32 * i;
      if (i < per && k < in_dim) acc = fmaf(wr[i], __ldg(a + k), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[(long long)b * total_out + j] = acc + bj;
  }
}

}  // namespace csd

extern "C" {

int csd_gn_chan_stats_bf16(const void* src, int c, int pitch, float* chan_sums, int batch, int hw,
                           csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(chan_sums != nullptr && batch >= 1 && hw >= 1, "gn_chan_stats: bad arguments");
  GnParams p;
  memset(&p, 0, sizeof(p));
  int threads;
  size_t smem;
  int st = gn_fill(p, src, c, pitch, nullptr, 0, 0, batch, hw, 1, &threads, &smem);
  if (st != CSD_OK) return st;
  p.sums0 = chan_sums;
  const size_t stat_smem = sizeof(float) * 2 * c * (size_t)(threads / (c / 8));
  gn_chan_stats_kernel<<<batch * p.slabs, threads, stat_smem, static_cast<cudaStream_t>(stream)>>>(p);
  CSD_LAUNCH_CHECK("gn_chan_stats_kernel");
  return CSD_OK;
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
  using namespace csd;
  CSD_REQUIRE(sums0 && gamma && beta && out && batch >= 1 && hw >= 1, "gn_apply: bad arguments");
  CSD_REQUIRE(src1 == nullptr || sums1 != nullptr, "gn_apply: second source without its channel sums");
  GnParams p;
  memset(&p, 0, sizeof(p));
  int threads;
  size_t smem;
  int st = gn_fill(p, src0, c0, pitch0, src1, c1, pitch1, batch, hw, groups, &threads, &smem);
  if (st != CSD_OK) return st;
  CSD_REQUIRE(out_pitch % 8 == 0 && out_pitch >= (p.v0 + p.v1) * 8, "gn_apply: out pitch %d too small", out_pitch);
  p.sums0 = const_cast<float*>(sums0);
  p.sums1 = const_cast<float*>(sums1);
  p.gamma = gamma; p.beta = beta;
  p.out = static_cast<bf16x8*>(out);
  p.out_pv = out_pitch / 8;
  p.eps = eps; p.silu = apply_silu;
  gn_apply_kernel<<<batch * p.slabs, threads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  CSD_LAUNCH_CHECK("gn_apply_kernel");
  return CSD_OK;
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

int csd_nchw_to_nhwc_bf16(const float* src0, int c0, const float* src1, int c1, void* out, int c_pad, int batch, int h,
                          int w, float scale, float shift, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(src0 && out && c0 >= 1, "nchw_to_nhwc: bad arguments");
  if (src1 == nullptr) c1 = 0;
  CSD_REQUIRE(c_pad % 8 == 0 && c_pad >= c0 + c1, "nchw_to_nhwc: c_pad=%d must be a multiple of 8 >= %d", c_pad, c0 + c1);
  const long long total = (long long)batch * h * w * (c_pad / 8);
  const int blocks = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)num_sms() * 16);
  nchw_to_nhwc_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src0, c0, src1, c1, static_cast<bf16x8*>(out),
                                                                           c_pad / 8, batch, (long long)h * w, scale, shift);
  CSD_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return CSD_OK;
}

int csd_nhwc_bf16_to_nchw(const void* src, int c_pitch, int c_off, int c_cnt, float* dst, int batch, int h, int w,
                          const float* row_scale, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(src && dst && c_cnt >= 1 && c_off >= 0 && c_off + c_cnt <= c_pitch, "nhwc_to_nchw: bad channel range");
  const long long total = (long long)batch * c_cnt * h * w;
  const int blocks = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)num_sms() * 16);
  nhwc_to_nchw_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), c_pitch, c_off, c_cnt, dst, batch, (long long)h * w, row_scale);
  CSD_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return CSD_OK;
}

int csd_softmax_rows_f32_bf16(const float* logits, int in_pitch, void* probs, int out_pitch, int64_t rows, int cols,
                              float scale, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(logits && probs && cols >= 1 && in_pitch >= cols && out_pitch >= cols, "softmax: bad arguments");
  if (rows == 0) return CSD_OK;
  const int blocks = (int)ceil_div_ll(rows, 8);
  softmax_rows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, in_pitch, static_cast<__nv_bfloat16*>(probs), out_pitch, rows, cols, scale);
  CSD_LAUNCH_CHECK("softmax_rows_kernel");
  return CSD_OK;
}

int csd_time_embedding_f32(const float* labels, int batch, int nf, int embedding_type, const float* fourier_w,
                           const float* w0, const float* b0, const float* w1, const float* b1, float* act_temb,
                           csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(labels && w0 && b0 && w1 && b1 && act_temb && batch >= 1 && nf >= 4, "time_embedding: bad arguments");
  CSD_REQUIRE(embedding_type == 0 || (embedding_type == 1 && fourier_w != nullptr), "time_embedding: bad embedding type");
  const size_t smem = sizeof(float) * ((embedding_type == 1 ? 2 * nf : nf) + 4 * nf);
  time_embedding_kernel<<<batch, 256, smem, static_cast<cudaStream_t>(stream)>>>(labels, nf, embedding_type, fourier_w, w0,
                                                                               b0, w1, b1, act_temb);
  CSD_LAUNCH_CHECK("time_embedding_kernel");
  return CSD_OK;
}

int csd_dense_rows_f32(const float* act_temb, const float* w, const float* bias, float* out, int batch, int in_dim,
                       int total_out, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(act_temb && w && out && batch >= 1 && in_dim >= 1 && in_dim <= 1024 && total_out >= 1,
              "dense_rows: bad arguments (in_dim <= 1024)");
  dense_rows_kernel<<<ceil_div(total_out, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(act_temb, w, bias, out, batch,
                                                                                         in_dim, total_out);
  CSD_LAUNCH_CHECK("dense_rows_kernel");
  return CSD_OK;
}

}  // extern "C"
