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
__global__ void __launch_bounds__(256)
ve_perturb_kernel(const float4* __restrict__ y, const float4* __restrict__ z, float4* __restrict__ out, long long n4,
                  const float* __restrict__ ys, const float* __restrict__ zs, float* __restrict__ os, int tail,
                  const float* __restrict__ sigma_tab, const int* __restrict__ step_idx) {
  const float sigma = sigma_tab[*step_idx];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = y[i], b = z[i];
    out[i] = make_float4(fmaf(b.x, sigma, a.x), fmaf(b.y, sigma, a.y), fmaf(b.z, sigma, a.z), fmaf(b.w, sigma, a.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < tail) os[threadIdx.x] = fmaf(zs[threadIdx.x], sigma, ys[threadIdx.x]);
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

__global__ void __launch_bounds__(256)
langevin_update_kernel(const float* __restrict__ x, const float* __restrict__ grad, const float* __restrict__ noise,
                       const float* __restrict__ norms, float* __restrict__ x_out, float* __restrict__ x_mean,
                       int batch, long long total, float snr, const float* __restrict__ alpha_tab,
                       const int* __restrict__ step_idx) {
  // batch-mean norms (correctors.py:72-73): every thread reduces the 2*batch floats (L1/L2 resident)
  float gsum = 0.f, zsum = 0.f;
  for (int b = 0; b < batch; ++b) {
    gsum += __ldg(norms + b);
    zsum += __ldg(norms + batch + b);
  }
  const float gn = gsum / batch, zn = zsum / batch;
  const float alpha = alpha_tab != nullptr ? alpha_tab[*step_idx] : 1.f;
  const float r = snr * zn / gn;
  const float step = r * r * 2.f * alpha;
  const float nz = sqrtf(step * 2.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float m = fmaf(step, grad[i], x[i]);
    x_mean[i] = m;
    x_out[i] = fmaf(nz, noise[i], m);
  }
}

__global__ void __launch_bounds__(256)
reverse_diffusion_kernel(const float* __restrict__ x, const float* __restrict__ score, const float* __restrict__ noise,
                         float* __restrict__ x_out, float* __restrict__ x_mean, long long n,
                         const float* __restrict__ f_tab, const float* __restrict__ g_tab, int pf,
                         const int* __restrict__ step_idx) {
  const int s = *step_idx;
  const float fc = f_tab != nullptr ? f_tab[s] : 0.f;
  const float g = g_tab[s];
  const float g2 = g * g * (pf ? 0.5f : 1.f);
  const float gz = pf ? 0.f : g;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xi = x[i];
    const float rev_f = fc * xi - g2 * score[i];
    const float m = xi - rev_f;
    x_mean[i] = m;
    x_out[i] = (noise != nullptr) ? fmaf(gz, noise[i], m) : m;
  }
}

__global__ void __launch_bounds__(256)
euler_maruyama_kernel(const float* __restrict__ x, const float* __restrict__ score, const float* __restrict__ noise,
                      float* __restrict__ x_out, float* __restrict__ x_mean, long long n,
                      const float* __restrict__ d_tab, const float* __restrict__ g_tab, float dt, int pf,
                      const int* __restrict__ step_idx) {
  const int s = *step_idx;
  const float dc = d_tab != nullptr ? d_tab[s] : 0.f;
  const float g = g_tab[s];
  const float g2 = g * g * (pf ? 0.5f : 1.f);
  const float gz = (pf ? 0.f : g) * sqrtf(-dt);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xi = x[i];
    const float drift = dc * xi - g2 * score[i];
    const float m = fmaf(drift, dt, xi);
    x_mean[i] = m;
    x_out[i] = (noise != nullptr) ? fmaf(gz, noise[i], m) : m;
  }
}

__global__ void step_advance_kernel(int* s) { *s += 1; }

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

int csd_ve_perturb_f32(const float* y, const float* z, float* y_pert, int64_t n, const float* sigma_tab,
                       const int* step_idx, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(y && z && y_pert && sigma_tab && step_idx, "ve_perturb: null pointer");
  CSD_REQUIRE(((uintptr_t)y & 15) == 0 && ((uintptr_t)z & 15) == 0 && ((uintptr_t)y_pert & 15) == 0,
              "ve_perturb: pointers must be 16-byte aligned");
  const long long n4 = n / 4;
  const int tail = (int)(n - n4 * 4);
  ve_perturb_kernel<<<ew_blocks(n4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(z), reinterpret_cast<float4*>(y_pert), n4,
      y + n4 * 4, z + n4 * 4, y_pert + n4 * 4, tail, sigma_tab, step_idx);
  CSD_LAUNCH_CHECK("ve_perturb_kernel");
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
                            const int* step_idx, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && grad && noise && norms && x_out && x_mean, "langevin_update: null pointer");
  CSD_REQUIRE(alpha_tab == nullptr || step_idx != nullptr, "langevin_update: alpha table without a step index");
  const long long total = (long long)batch * per_sample;
  langevin_update_kernel<<<ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, grad, noise, norms, x_out, x_mean, batch, total, snr, alpha_tab, step_idx);
  CSD_LAUNCH_CHECK("langevin_update_kernel");
  return CSD_OK;
}

int csd_reverse_diffusion_update_f32(const float* x, const float* score, const float* noise, float* x_out,
                                     float* x_mean, int64_t n, const float* f_coef_tab, const float* g_tab,
                                     int probability_flow, const int* step_idx, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && score && x_out && x_mean && g_tab && step_idx, "reverse_diffusion_update: null pointer");
  CSD_REQUIRE(noise != nullptr || probability_flow, "reverse_diffusion_update: noise required unless probability flow");
  reverse_diffusion_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, score, noise, x_out, x_mean, n, f_coef_tab, g_tab, probability_flow, step_idx);
  CSD_LAUNCH_CHECK("reverse_diffusion_kernel");
  return CSD_OK;
}

int csd_euler_maruyama_update_f32(const float* x, const float* score, const float* noise, float* x_out, float* x_mean,
                                  int64_t n, const float* d_coef_tab, const float* g_tab, float dt,
                                  int probability_flow, const int* step_idx, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(x && score && x_out && x_mean && g_tab && step_idx, "euler_maruyama_update: null pointer");
  CSD_REQUIRE(noise != nullptr || probability_flow, "euler_maruyama_update: noise required unless probability flow");
  CSD_REQUIRE(dt < 0.f, "euler_maruyama_update: dt must be negative (reverse time)");
  euler_maruyama_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, score, noise, x_out, x_mean, n, d_coef_tab, g_tab, dt, probability_flow, step_idx);
  CSD_LAUNCH_CHECK("euler_maruyama_kernel");
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
