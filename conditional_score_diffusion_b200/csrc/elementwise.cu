// HBM-bound elementwise kernels of the hot path: fused_bias_act (the reference's second native op)
// and the predictor-corrector update kernels.
//
// References: op/fused_bias_act_kernel.cu:18-99; sampling/correctors.py:51-108;
// sampling/predictors.py:52-102; sampling/conditional.py:104-110; sde_lib.py:49-63,87-92,349-360.
// The reference runs each update as 5-12 separate ATen kernels plus host-built scalars; here every
// update is ONE pass over the tensors (16-byte vector loads/stores, grid sized to the SM count) and
// the per-step scalars come from device tables indexed by a device-resident step counter, so a
// captured CUDA graph of a whole PC step replays without host work.
#include "common.cuh"
#include "../../include/csd_b200.h"

#include <algorithm>
#include <initializer_list>

namespace csd {

static inline int ew_blocks(long long n_vec) {
  long long b = ceil_div_ll(n_vec, 256);
  long long cap = (long long)num_sms() * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// ---- fused_bias_act ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fused_bias_act_kernel(const float* __restrict__ x, const float* __restrict__ bias, const float* __restrict__ ref,
                      float* __restrict__ y, long long n, int size_b, long long step_b, int act, int grad,
                      float alpha, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    if (bias != nullptr) v += __ldg(bias + (i / step_b) % size_b);
    const float r = ref != nullptr ? ref[i] : 0.f;
    float o;
    if (act == 3) {                       // leaky relu (fused_bias_act_kernel.cu:36-45)
      if (grad == 0) o = v > 0.f ? v : v * alpha;
      else if (grad == 1) o = r > 0.f ? v : v * alpha;
      else o = 0.f;
    } else {                              // linear
      o = grad == 2 ? 0.f : v;
    }
    y[i] = o * scale;
  }
}

// ---- PC updates ------------------------------------------------------------------------------------
// Per-step / per-sample scalar lookup. Tables are [n_steps] (sample_stride = 0: one value per step,
// shared by the batch) or [n_steps, batch] (sample_stride = 1). step_idx may be null (step 0), which
// lets the class-based predictors / correctors pass plain per-sample [batch] arrays.
struct CoefRef {
  const float* tab;
  const int* step_idx;
  int sample_stride;
  int batch;
  __device__ __forceinline__ float at(int b, float dflt) const {
    if (tab == nullptr) return dflt;
    const int s = step_idx != nullptr ? *step_idx : 0;
    return tab[(long long)s * (sample_stride ? batch : 1) + (long long)b * sample_stride];
  }
};

// All update kernels use grid = (blocks per sample, batch): the sample index is blockIdx.y, so the
// per-sample coefficients are CTA constants and the inner loop is pure 16-byte streaming.
template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<1> { using type = float; };

template <int VEC, typename F>
__device__ __forceinline__ void map3(const float* a, const float* b, const float* c, float* o0, float* o1,
                                     long long per_sample, F f) {
  // o0[i], o1[i] = f(a[i], b[i], c[i]); any of b, c, o1 may be null
  const long long base = (long long)blockIdx.y * per_sample;
  const long long nvec = per_sample / VEC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float va[VEC], vb[VEC], vc[VEC], r0[VEC], r1[VEC];
    using V = typename VecT<VEC>::type;
    *reinterpret_cast<V*>(va) = *reinterpret_cast<const V*>(a + base + i * VEC);
    if (b != nullptr) *reinterpret_cast<V*>(vb) = *reinterpret_cast<const V*>(b + base + i * VEC);
    if (c != nullptr) *reinterpret_cast<V*>(vc) = *reinterpret_cast<const V*>(c + base + i * VEC);
#pragma unroll
    for (int k = 0; k < VEC; ++k) f(va[k], b != nullptr ? vb[k] : 0.f, c != nullptr ? vc[k] : 0.f, r0[k], r1[k]);
    *reinterpret_cast<V*>(o0 + base + i * VEC) = *reinterpret_cast<V*>(r0);
    if (o1 != nullptr) *reinterpret_cast<V*>(o1 + base + i * VEC) = *reinterpret_cast<V*>(r1);
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
ve_perturb_kernel(const float* __restrict__ y, const float* __restrict__ z, float* __restrict__ out,
                  long long per_sample, CoefRef sig) {
  const float sigma = sig.at(blockIdx.y, 0.f);
  map3<VEC>(y, z, nullptr, out, nullptr, per_sample,
            [sigma](float a, float b, float, float& r0, float&) { r0 = fmaf(b, sigma, a); });
}

// Forward perturbation of the denoising score-matching losses (losses.py:126-133,190-192,218-220):
// out = mean_coef[b] * x + std[b] * z (mean_coef null = 1: VE SDEs).
template <int VEC>
__global__ void __launch_bounds__(256)
sde_perturb_kernel(const float* __restrict__ x, const float* __restrict__ z, float* __restrict__ out,
                   long long per_sample, const float* __restrict__ mean_coef, const float* __restrict__ stdv) {
  const float m = mean_coef != nullptr ? mean_coef[blockIdx.y] : 1.f;
  const float sd = stdv[blockIdx.y];
  map3<VEC>(x, z, nullptr, out, nullptr, per_sample,
            [m, sd](float a, float b, float, float& r0, float&) { r0 = fmaf(b, sd, m * a); });
}

// Per-sample denoising score-matching residual: losses[b] += w[b] * sum_i (a[b]*score_i + c[b]*z_i)^2
// (losses.py:139-145 / 197-203 / 223-229: a = 1, c = 1/std, w = g^2 * reduce factor with likelihood weighting;
// a = std, c = 1 without). One CTA row per sample slab, partial sums combined with atomics.
__global__ void __launch_bounds__(256)
dsm_loss_kernel(const float* __restrict__ score, const float* __restrict__ z, const float* __restrict__ a,
                const float* __restrict__ c, const float* __restrict__ w, float* __restrict__ losses,
                long long per_sample, int slabs) {
  const int b = blockIdx.x / slabs, slab = blockIdx.x % slabs;
  const long long chunk = ceil_div_ll(per_sample, slabs);
  const long long lo = slab * chunk, hi = min(per_sample, lo + chunk);
  const float* sp = score + (long long)b * per_sample;
  const float* zp = z + (long long)b * per_sample;
  const float ab = a[b], cb = c[b];
  float acc = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float r = fmaf(ab, sp[i], cb * zp[i]);
    acc = fmaf(r, r, acc);
  }
  acc = warp_sum(acc);
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(losses + b, t * w[b]);
  }
}

// One CTA row per sample slab; partial sums of squares are combined with atomics into sq[2*batch].
__global__ void __launch_bounds__(256)
sumsq_pair_kernel(const float* __restrict__ g, const float* __restrict__ z, float* __restrict__ sq, int batch,
                  long long per_sample, int slabs) {
  const int b = blockIdx.x / slabs, slab = blockIdx.x % slabs;
  const long long chunk = ceil_div_ll(per_sample, slabs);
  const long long lo = slab * chunk, hi = min(per_sample, lo + chunk);
  const float* gp = g + (long long)b * per_sample;
  const float* zp = z + (long long)b * per_sample;
  float sg = 0.f, sz = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float a = gp[i], c = zp[i];
    sg = fmaf(a, a, sg);
    sz = fmaf(c, c, sz);
  }
  sg = warp_sum(sg);
  sz = warp_sum(sz);
  __shared__ float red[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = sg; red[1][warp] = sz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; c += red[1][w]; }
    atomicAdd(sq + b, a);
    atomicAdd(sq + batch + b, c);
  }
}

__global__ void sqrt_inplace_kernel(float* v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = sqrtf(v[i]);
}

template <int VEC>
__global__ void __launch_bounds__(256)
langevin_update_kernel(const float* x, const float* __restrict__ grad, const float* __restrict__ noise,
                       const float* __restrict__ norms, float* x_out, float* __restrict__ x_mean,
                       int batch, long long per_sample, float snr, CoefRef alpha_ref) {
  // batch-mean norms (correctors.py:72-73): every thread reduces the 2*batch floats (L1/L2 resident)
  float gsum = 0.f, zsum = 0.f;
  for (int b = 0; b < batch; ++b) {
    gsum += __ldg(norms + b);
    zsum += __ldg(norms + batch + b);
  }
  const float gn = gsum / batch, zn = zsum / batch;
  const float r = snr * zn / gn;
  const float step = r * r * 2.f * alpha_ref.at(blockIdx.y, 1.f);
  const float nz = sqrtf(step * 2.f);
  map3<VEC>(x, grad, noise, x_out, x_mean, per_sample, [step, nz](float xi, float g, float z, float& xo, float& xm) {
    xm = fmaf(step, g, xi);
    xo = fmaf(nz, z, xm);
  });
}

template <int VEC>
__global__ void __launch_bounds__(256)
reverse_diffusion_kernel(const float* x, const float* __restrict__ score, const float* __restrict__ noise,
                         float* x_out, float* __restrict__ x_mean, long long per_sample, CoefRef f_ref,
                         CoefRef g_ref, int pf) {
  const float fc = f_ref.at(blockIdx.y, 0.f);
  const float g = g_ref.at(blockIdx.y, 0.f);
  const float g2 = g * g * (pf ? 0.5f : 1.f);
  const float gz = pf ? 0.f : g;
  map3<VEC>(x, score, noise, x_out, x_mean, per_sample, [fc, g2, gz](float xi, float sc, float z, float& xo, float& xm) {
    const float rev_f = fc * xi - g2 * sc;
    xm = xi - rev_f;
    xo = fmaf(gz, z, xm);
  });
}

template <int VEC>
__global__ void __launch_bounds__(256)
euler_maruyama_kernel(const float* x, const float* __restrict__ score, const float* __restrict__ noise,
                      float* x_out, float* __restrict__ x_mean, long long per_sample, CoefRef d_ref,
                      CoefRef g_ref, float dt, int pf) {
  const float dc = d_ref.at(blockIdx.y, 0.f);
  const float g = g_ref.at(blockIdx.y, 0.f);
  const float g2 = g * g * (pf ? 0.5f : 1.f);
  const float gz = (pf ? 0.f : g) * sqrtf(-dt);
  map3<VEC>(x, score, noise, x_out, x_mean, per_sample, [dc, g2, gz, dt](float xi, float sc, float z, float& xo, float& xm) {
    const float drift = dc * xi - g2 * sc;
    xm = fmaf(drift, dt, xi);
    xo = fmaf(gz, z, xm);
  });
}

// Inpainting projection after a predictor / corrector update (sampling/unconditional.py:266-277):
// masked = mean_coef * data + std * z;  x_out = x (1 - m) + masked m;  x_mean = x_out (1 - m) + mean_coef * data * m
// (the reference builds x_mean from the already merged x: kept as is).
__global__ void __launch_bounds__(256)
inpaint_merge_kernel(const float* x, const float* __restrict__ data, const float* __restrict__ z,
                     const float* __restrict__ mask, float* x_out, float* __restrict__ x_mean, long long per_sample,
                     const float* __restrict__ mean_coef, const float* __restrict__ stdv) {
  const int b = blockIdx.y;
  const float mc = mean_coef != nullptr ? mean_coef[b] : 1.f;
  const float sd = stdv[b];
  const long long base = (long long)b * per_sample;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x) {
    const float m = mask[base + i];
    const float dm = mc * data[base + i];
    const float xo = x[base + i] * (1.f - m) + fmaf(sd, z[base + i], dm) * m;
    x_out[base + i] = xo;
    x_mean[base + i] = xo * (1.f - m) + dm * m;
  }
}

// ---- device-resident Runge-Kutta pieces (probability-flow ODE sampler / likelihood) ------------
// k is a stack of stage derivatives [stages, n]; out = y + h * sum_i coef[i] * k[i] for i < ns.
struct RkCoefs { float c[8]; };

__global__ void __launch_bounds__(256)
rk_combine_kernel(const float* __restrict__ y, const float* __restrict__ k, long long n, int ns, RkCoefs coef, float h,
                  float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < ns; ++s) acc = fmaf(coef.c[s], k[(long long)s * n + i], acc);
    out[i] = fmaf(h, acc, y[i]);
  }
}

// out[0] += sum_i ( h * sum_s e[s] k[s][i] / (atol + max(|y_i|, |y2_i|) * rtol) )^2
__global__ void __launch_bounds__(256)
rk_error_sumsq_kernel(const float* __restrict__ k, long long n, int ns, RkCoefs e, float h, const float* __restrict__ y,
                      const float* __restrict__ y2, float atol, float rtol, float* out) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float err = 0.f;
    for (int s = 0; s < ns; ++s) err = fmaf(e.c[s], k[(long long)s * n + i], err);
    const float sc = fmaf(fmaxf(fabsf(y[i]), fabsf(y2[i])), rtol, atol);
    const float r = h * err / sc;
    acc = fmaf(r, r, acc);
  }
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

// ---- fused optimizer step over the flat parameter buffer --------------------------------------------------------
// out[0] += sum x^2 (block partials combined with one atomic per block).
__global__ void __launch_bounds__(256) sumsq_f32_kernel(const float* __restrict__ x, long long n, float* out) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc = fmaf(x[i], x[i], acc);
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(out, t);
  }
}

// clip_grad_norm_ + torch.optim.Adam (L2 weight decay, no amsgrad) + ExponentialMovingAverage.update in one pass:
//   g *= min(1, max_norm / (||g|| + 1e-6));  g += wd p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps);  ema -= (1 - decay) (ema - p)
__global__ void __launch_bounds__(256)
fused_adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                      float* __restrict__ ema, long long n, float lr, float b1, float b2, float eps, float wd, float bc1,
                      float rsqrt_bc2, float max_norm, const float* __restrict__ gnorm_sq, float one_minus_decay) {
  float clip = 1.f;
  if (max_norm >= 0.f) clip = fminf(1.f, max_norm / (sqrtf(*gnorm_sq) + 1e-6f));
  const float step = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * clip);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    pi -= step * mi / fmaf(sqrtf(vi), rsqrt_bc2, eps);
    p[i] = pi;
    if (ema != nullptr) {
      const float e = ema[i];
      ema[i] = e - one_minus_decay * (e - pi);
    }
  }
}

static inline dim3 ps_grid(int batch, long long per_sample, int vec) {
  long long bx = ceil_div_ll(per_sample / vec, 256);
  long long cap = std::max<long long>(1, ceil_div_ll((long long)num_sms() * 8, batch));
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)batch, 1);
}

static inline bool vec4_ok(long long per_sample, std::initializer_list<const void*> ptrs) {
  if (per_sample % 4 != 0) return false;
  for (const void* p : ptrs)
    if (p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15) != 0) return false;
  return true;
}

__global__ void step_advance_kernel(int* s) { *s += 1; }

__global__ void broadcast_table_kernel(float* dst, int n, CoefRef ref) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = ref.at(i, 0.f);
}

}  // namespace csd

extern "C" {

int csd_fused_bias_act_f32(const float* x, const float* bias, const float* refer, float* y, int64_t n, int size_b,
                           int64_t step_b, int act, int grad, float alpha, float scale, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && y, "fused_bias_act: null pointer");
  CSD_REQUIRE(act == 1 || act == 3, "fused_bias_act: act %d (reference kernel has 1 = linear, 3 = lrelu)", act);
  CSD_REQUIRE(grad >= 0 && grad <= 2, "fused_bias_act: grad %d", grad);
  CSD_REQUIRE(bias == nullptr || (size_b >= 1 && step_b >= 1), "fused_bias_act: bad bias geometry");
  if (n == 0) return CSD_OK;
  fused_bias_act_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, bias, refer, y, n, size_b, step_b,
                                                                                    act, grad, alpha, scale);
  CSD_LAUNCH_CHECK("fused_bias_act_kernel");
  return CSD_OK;
}

int csd_ve_perturb_f32(const float* y, const float* z, float* y_pert, int batch, int64_t per_sample,
                       const float* sigma_tab, const int* step_idx, int sample_stride, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(y && z && y_pert && sigma_tab, "ve_perturb: null pointer");
  CSD_REQUIRE(batch >= 1 && batch <= 65535 && per_sample >= 1, "ve_perturb: bad batch / per_sample");
  CoefRef sig{sigma_tab, step_idx, sample_stride, batch};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec4_ok(per_sample, {y, z, y_pert}))
    ve_perturb_kernel<4><<<ps_grid(batch, per_sample, 4), 256, 0, st>>>(y, z, y_pert, per_sample, sig);
  else
    ve_perturb_kernel<1><<<ps_grid(batch, per_sample, 1), 256, 0, st>>>(y, z, y_pert, per_sample, sig);
  CSD_LAUNCH_CHECK("ve_perturb_kernel");
  return CSD_OK;
}

int csd_sde_perturb_f32(const float* x, const float* z, float* out, int batch, int64_t per_sample,
                        const float* mean_coef, const float* std_dev, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && z && out && std_dev, "sde_perturb: null pointer");
  CSD_REQUIRE(batch >= 1 && batch <= 65535 && per_sample >= 1, "sde_perturb: bad batch / per_sample");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec4_ok(per_sample, {x, z, out}))
    sde_perturb_kernel<4><<<ps_grid(batch, per_sample, 4), 256, 0, st>>>(x, z, out, per_sample, mean_coef, std_dev);
  else
    sde_perturb_kernel<1><<<ps_grid(batch, per_sample, 1), 256, 0, st>>>(x, z, out, per_sample, mean_coef, std_dev);
  CSD_LAUNCH_CHECK("sde_perturb_kernel");
  return CSD_OK;
}

int csd_rk_combine_f32(const float* y, const float* k_stack, int64_t n, int stages, const float* coef_host, float h,
                       float* out, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(y && k_stack && out && coef_host && n >= 1 && stages >= 1 && stages <= 8, "rk_combine: bad arguments");
  RkCoefs c;
  for (int i = 0; i < 8; ++i) c.c[i] = i < stages ? coef_host[i] : 0.f;
  rk_combine_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, k_stack, n, stages, c, h, out);
  CSD_LAUNCH_CHECK("rk_combine_kernel");
  return CSD_OK;
}

int csd_rk_error_sumsq_f32(const float* k_stack, int64_t n, int stages, const float* e_host, float h, const float* y,
                           const float* y2, float atol, float rtol, float* out, csd_stream_t stream_) {
  using namespace csd;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CSD_REQUIRE(k_stack && y && y2 && out && e_host && n >= 1 && stages >= 1 && stages <= 8, "rk_error_sumsq: bad arguments");
  RkCoefs e;
  for (int i = 0; i < 8; ++i) e.c[i] = i < stages ? e_host[i] : 0.f;
  CSD_CUDA(cudaMemsetAsync(out, 0, sizeof(float), stream));
  rk_error_sumsq_kernel<<<ew_blocks(n), 256, 0, stream>>>(k_stack, n, stages, e, h, y, y2, atol, rtol, out);
  CSD_LAUNCH_CHECK("rk_error_sumsq_kernel");
  return CSD_OK;
}

int csd_sumsq_f32(const float* x, int64_t n, float* out, csd_stream_t stream_) {
  using namespace csd;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CSD_REQUIRE(x && out && n >= 1, "sumsq: bad arguments");
  CSD_CUDA(cudaMemsetAsync(out, 0, sizeof(float), stream));
  sumsq_f32_kernel<<<ew_blocks(n), 256, 0, stream>>>(x, n, out);
  CSD_LAUNCH_CHECK("sumsq_f32_kernel");
  return CSD_OK;
}

int csd_fused_adam_ema_f32(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr, float beta1,
                           float beta2, float eps, float weight_decay, float bias_corr1, float bias_corr2, float max_norm,
                           const float* gnorm_sq, float ema_decay, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(p && g && m && v && n >= 1, "fused_adam_ema: null pointer");
  CSD_REQUIRE(max_norm < 0.f || gnorm_sq != nullptr, "fused_adam_ema: clipping needs the squared gradient norm");
  CSD_REQUIRE(bias_corr1 > 0.f && bias_corr2 > 0.f, "fused_adam_ema: bias corrections must be positive");
  fused_adam_ema_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, ema, n, lr, beta1, beta2, eps, weight_decay, bias_corr1, rsqrtf(bias_corr2), max_norm, gnorm_sq,
      1.f - ema_decay);
  CSD_LAUNCH_CHECK("fused_adam_ema_kernel");
  return CSD_OK;
}

int csd_inpaint_merge_f32(const float* x, const float* data, const float* z, const float* mask, float* x_out, float* x_mean,
                          int batch, int64_t per_sample, const float* mean_coef, const float* std_dev, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && data && z && mask && x_out && x_mean && std_dev, "inpaint_merge: null pointer");
  CSD_REQUIRE(batch >= 1 && batch <= 65535 && per_sample >= 1, "inpaint_merge: bad batch / per_sample");
  inpaint_merge_kernel<<<ps_grid(batch, per_sample, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, data, z, mask, x_out, x_mean, per_sample, mean_coef, std_dev);
  CSD_LAUNCH_CHECK("inpaint_merge_kernel");
  return CSD_OK;
}

int csd_dsm_loss_f32(const float* score, const float* z, const float* a, const float* c, const float* w, float* losses,
                     int batch, int64_t per_sample, csd_stream_t stream_) {
  using namespace csd;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CSD_REQUIRE(score && z && a && c && w && losses && batch >= 1 && per_sample >= 1, "dsm_loss: bad arguments");
  int slabs = (int)std::max<long long>(1, std::min<long long>(ceil_div_ll(per_sample, 8192),
                                                              ceil_div_ll((long long)num_sms() * 4, batch)));
  dsm_loss_kernel<<<batch * slabs, 256, 0, stream>>>(score, z, a, c, w, losses, per_sample, slabs);
  CSD_LAUNCH_CHECK("dsm_loss_kernel");
  return CSD_OK;
}

int csd_langevin_norms_f32(const float* grad, const float* noise, float* norms, int batch, int64_t per_sample,
                           csd_stream_t stream_) {
  using namespace csd;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CSD_REQUIRE(grad && noise && norms && batch >= 1 && per_sample >= 1, "langevin_norms: bad arguments");
  CSD_CUDA(cudaMemsetAsync(norms, 0, sizeof(float) * 2 * batch, stream));
  int slabs = (int)std::max<long long>(1, std::min<long long>(ceil_div_ll(per_sample, 8192),
                                                              ceil_div_ll((long long)num_sms() * 4, batch)));
  sumsq_pair_kernel<<<batch * slabs, 256, 0, stream>>>(grad, noise, norms, batch, per_sample, slabs);
  CSD_LAUNCH_CHECK("sumsq_pair_kernel");
  sqrt_inplace_kernel<<<ceil_div(2 * batch, 128), 128, 0, stream>>>(norms, 2 * batch);
  CSD_LAUNCH_CHECK("sqrt_inplace_kernel");
  return CSD_OK;
}

int csd_langevin_update_f32(const float* x, const float* grad, const float* noise, const float* norms, float* x_out,
                            float* x_mean, int batch, int64_t per_sample, float snr, const float* alpha_tab,
                            const int* step_idx, int sample_stride, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && grad && noise && norms && x_out && x_mean, "langevin_update: null pointer");
  CSD_REQUIRE(batch >= 1 && batch <= 65535 && per_sample >= 1, "langevin_update: bad batch / per_sample");
  CoefRef a{alpha_tab, step_idx, sample_stride, batch};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec4_ok(per_sample, {x, grad, noise, x_out, x_mean}))
    langevin_update_kernel<4><<<ps_grid(batch, per_sample, 4), 256, 0, st>>>(x, grad, noise, norms, x_out, x_mean, batch,
                                                                            per_sample, snr, a);
  else
    langevin_update_kernel<1><<<ps_grid(batch, per_sample, 1), 256, 0, st>>>(x, grad, noise, norms, x_out, x_mean, batch,
                                                                            per_sample, snr, a);
  CSD_LAUNCH_CHECK("langevin_update_kernel");
  return CSD_OK;
}

int csd_reverse_diffusion_update_f32(const float* x, const float* score, const float* noise, float* x_out,
                                     float* x_mean, int batch, int64_t per_sample, const float* f_coef_tab,
                                     const float* g_tab, int probability_flow, const int* step_idx,
                                     int sample_stride, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && score && x_out && x_mean && g_tab, "reverse_diffusion_update: null pointer");
  CSD_REQUIRE(noise != nullptr || probability_flow, "reverse_diffusion_update: noise required unless probability flow");
  CSD_REQUIRE(batch >= 1 && batch <= 65535 && per_sample >= 1, "reverse_diffusion_update: bad batch / per_sample");
  CoefRef f{f_coef_tab, step_idx, sample_stride, batch}, g{g_tab, step_idx, sample_stride, batch};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec4_ok(per_sample, {x, score, noise, x_out, x_mean}))
    reverse_diffusion_kernel<4><<<ps_grid(batch, per_sample, 4), 256, 0, st>>>(x, score, noise, x_out, x_mean, per_sample,
                                                                              f, g, probability_flow);
  else
    reverse_diffusion_kernel<1><<<ps_grid(batch, per_sample, 1), 256, 0, st>>>(x, score, noise, x_out, x_mean, per_sample,
                                                                              f, g, probability_flow);
  CSD_LAUNCH_CHECK("reverse_diffusion_kernel");
  return CSD_OK;
}

int csd_euler_maruyama_update_f32(const float* x, const float* score, const float* noise, float* x_out, float* x_mean,
                                  int batch, int64_t per_sample, const float* d_coef_tab, const float* g_tab, float dt,
                                  int probability_flow, const int* step_idx, int sample_stride, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && score && x_out && x_mean && g_tab, "euler_maruyama_update: null pointer");
  CSD_REQUIRE(noise != nullptr || probability_flow, "euler_maruyama_update: noise required unless probability flow");
  CSD_REQUIRE(dt < 0.f, "euler_maruyama_update: dt must be negative (reverse time)");
  CSD_REQUIRE(batch >= 1 && batch <= 65535 && per_sample >= 1, "euler_maruyama_update: bad batch / per_sample");
  CoefRef d{d_coef_tab, step_idx, sample_stride, batch}, g{g_tab, step_idx, sample_stride, batch};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec4_ok(per_sample, {x, score, noise, x_out, x_mean}))
    euler_maruyama_kernel<4><<<ps_grid(batch, per_sample, 4), 256, 0, st>>>(x, score, noise, x_out, x_mean, per_sample, d,
                                                                           g, dt, probability_flow);
  else
    euler_maruyama_kernel<1><<<ps_grid(batch, per_sample, 1), 256, 0, st>>>(x, score, noise, x_out, x_mean, per_sample, d,
                                                                           g, dt, probability_flow);
  CSD_LAUNCH_CHECK("euler_maruyama_kernel");
  return CSD_OK;
}

int csd_broadcast_table_f32(float* dst, int n, const float* tab, const int* step_idx, int sample_stride,
                            csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(dst && tab && n >= 1, "broadcast_table: bad arguments");
  CoefRef r{tab, step_idx, sample_stride, n};
  broadcast_table_kernel<<<ceil_div(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(dst, n, r);
  CSD_LAUNCH_CHECK("broadcast_table_kernel");
  return CSD_OK;
}

int csd_step_advance(int* step_idx, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(step_idx, "step_advance: null pointer");
  step_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(step_idx);
  CSD_LAUNCH_CHECK("step_advance_kernel");
  return CSD_OK;
}

}  // extern "C"
