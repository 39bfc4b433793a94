/* libcsd_b200 — C ABI of the B200 (sm_100a) score-network / PC-sampler hot path.
 *
 * Drop-in boundary for GBATZOLIS/conditional_score_diffusion. The reference reaches native
 * code through two pybind11 modules built at import time:
 *   op/upfirdn2d.cpp:12-23      upfirdn2d(Tensor input, Tensor kernel, up_x, up_y, down_x, down_y,
 *                                         pad_x0, pad_x1, pad_y0, pad_y1) -> Tensor
 *   op/fused_bias_act.cpp:11-21 fused_bias_act(Tensor input, Tensor bias, Tensor refer, act, grad,
 *                                              alpha, scale) -> Tensor
 * and everything else on the hot path is ATen/cuDNN eager PyTorch. This library
 * replaces both native modules and the eager op sequences of the score network and the
 * predictor-corrector update, behind plain C entry points:
 *
 *   - every pointer is a DEVICE pointer unless the name says host; no torch / ATen types;
 *   - outputs are caller-allocated; nothing here allocates device memory except
 *     csd_program_create (tensor maps and the op table, host memory only);
 *   - every launch goes to the cudaStream_t passed by the caller (pass torch's current stream);
 *   - the return value is 0 (CSD_OK) or a non-zero status; csd_last_error() gives the text
 *     (thread-local). Launch errors are checked after every launch (the reference's ops do not:
 *     op/upfirdn2d_kernel.cu:209-369 has no cudaGetLastError).
 *
 * Internal activation layout of the score network: NHWC, bf16, channel pitch a multiple of 8
 * (16 bytes). The module-surface ops (upfirdn2d, fused_bias_act, PC updates) are NCHW fp32
 * like the reference's tensors.
 */
#ifndef CSD_B200_H_
#define CSD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSD_ABI_VERSION 1

typedef void* csd_stream_t; /* cudaStream_t */

/* ---- runtime ------------------------------------------------------------------------------ */
const char* csd_last_error(void);
int csd_abi_version(void);
int csd_device_sm_count(int* out);
/* Kernels launched by this library in this process so far (host-side count of <<<>>> launches; a
 * captured CUDA graph replays them without passing here again). */
long long csd_launch_count(void);

/* ---- upfirdn2d: replaces op/upfirdn2d.cpp:12-23 + op/upfirdn2d_kernel.cu:209-369 -------------
 * input  [planes, in_h, in_w] fp32 (planes = N*C; the reference views it as [major,H,W,minor=1],
 *        op/upfirdn2d.py:99), kernel [kh, kw] fp32 (device), output [planes, out_h, out_w] fp32,
 *        out = ((in*up + pad0 + pad1 - k) / down) + 1 (op/upfirdn2d.py:104-105).
 * The backward pass is the same entry with up/down swapped, the flipped kernel and g_pad
 * (op/upfirdn2d.py:25-44), so there is no separate backward symbol.                            */
int csd_upfirdn2d_f32(const float* input, const float* kernel, float* output, int64_t planes, int in_h,
                      int in_w, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0,
                      int pad_x1, int pad_y0, int pad_y1, csd_stream_t stream);
int csd_upfirdn2d_out_size(int in_h, int in_w, int kh, int kw, int up_x, int up_y, int down_x,
                           int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1, int* out_h,
                           int* out_w);

/* ---- fused_bias_act: replaces op/fused_bias_act.cpp:11-21 + fused_bias_act_kernel.cu:18-99 ---
 * y = act(x + b[(i / step_b) % size_b]) * scale, act 1 = linear, 3 = leaky-relu(alpha);
 * grad 0 = forward, 1 = first derivative w.r.t. x using refer's sign, 2 = second derivative (0
 * for these activations). bias / refer may be NULL (size_b = 0 / no reference).                */
int csd_fused_bias_act_f32(const float* x, const float* bias, const float* refer, float* y, int64_t n,
                           int size_b, int64_t step_b, int act, int grad, float alpha, float scale,
                           csd_stream_t stream);

/* ---- predictor-corrector updates (sampling/predictors.py:79-102, correctors.py:51-108,
 *      sampling/conditional.py:104-110). All tensors [batch, per_sample] fp32, contiguous.
 *      Per-step scalars live in DEVICE tables indexed by *step_idx (a device int) so a captured
 *      CUDA graph can be replayed for every step without host involvement.                     */

/* Scalar lookup convention shared by the entry points below: a coefficient table is either
 * [n_steps] (sample_stride = 0: one value per step, shared by the batch) or [n_steps, batch]
 * (sample_stride = 1); step_idx is a DEVICE int (NULL = step 0, i.e. a plain [batch] array).      */

/* y_pert[b] = y[b] + z[b] * sigma(b)  (VESDE.marginal_prob, sde_lib.py:316-321;
 * sampling/conditional.py:107-108). per_sample must be a multiple of 4.                          */
int csd_ve_perturb_f32(const float* y, const float* z, float* y_pert, int batch, int64_t per_sample,
                       const float* sigma_tab, const int* step_idx, int sample_stride, csd_stream_t stream);

/* norms[0..batch) = ||grad_b||_2, norms[batch..2*batch) = ||noise_b||_2 (correctors.py:72-73). One launch; a
 * thread-block cluster per sample sums in a fixed order (bitwise reproducible: these norms set the Langevin step size
 * of the whole batch).                                                                            */
int csd_langevin_norms_f32(const float* grad, const float* noise, float* norms, int batch,
                           int64_t per_sample, csd_stream_t stream);

/* step = 2*alpha(b)*(snr*mean(noise_norm)/mean(grad_norm))^2; x_mean = x + step*grad;
 * x_out = x_mean + sqrt(2*step)*noise (correctors.py:74-76). alpha_tab NULL = 1 (VE).
 * x_out may alias x.                                                                           */
int csd_langevin_update_f32(const float* x, const float* grad, const float* noise, const float* norms,
                            float* x_out, float* x_mean, int batch, int64_t per_sample, float snr,
                            const float* alpha_tab, const int* step_idx, int sample_stride,
                            csd_stream_t stream);

/* Reverse-diffusion predictor (predictors.py:79-102 with sde_lib.py:87-92,186-194,349-360):
 * rev_f = f_coef(b)*x - g(b)^2*score*(pf ? .5 : 1); x_mean = x - rev_f; x_out = x_mean + g*noise
 * (no noise term if probability_flow). f_coef_tab NULL = 0 (VE). x_out may alias x.              */
int csd_reverse_diffusion_update_f32(const float* x, const float* score, const float* noise, float* x_out,
                                     float* x_mean, int batch, int64_t per_sample, const float* f_coef_tab,
                                     const float* g_tab, int probability_flow, const int* step_idx,
                                     int sample_stride, csd_stream_t stream);

/* Euler-Maruyama predictor (predictors.py:52-77): drift = d_coef(b)*x - g(b)^2*score*(pf ? .5 : 1);
 * x_mean = x + drift*dt; x_out = x_mean + g*sqrt(-dt)*noise, dt = -1/N.                          */
int csd_euler_maruyama_update_f32(const float* x, const float* score, const float* noise, float* x_out,
                                  float* x_mean, int batch, int64_t per_sample, const float* d_coef_tab,
                                  const float* g_tab, float dt, int probability_flow, const int* step_idx,
                                  int sample_stride, csd_stream_t stream);

/* ---- denoising score-matching losses (losses.py:99-234) ---------------------------------------------
 * out[b] = mean_coef[b] * x[b] + std_dev[b] * z[b]: SDE.marginal_prob mean + std * z (losses.py:126-133,190-192,
 * 218-220). mean_coef NULL = 1 (VE SDEs). All [batch] arrays are device pointers.                      */
int csd_sde_perturb_f32(const float* x, const float* z, float* out, int batch, int64_t per_sample,
                        const float* mean_coef, const float* std_dev, csd_stream_t stream);

/* losses[b] += w[b] * sum_i (a[b] * score[b,i] + c[b] * z[b,i])^2 (losses.py:139-145,197-203,223-229:
 * likelihood weighting a = 1, c = 1/std, w = g(t)^2 * (1/per_sample or 1/2); otherwise a = std, c = 1).
 * The caller zeroes `losses` and may accumulate several terms (the x and y parts of the CMDE loss).   */
int csd_dsm_loss_f32(const float* score, const float* z, const float* a, const float* c, const float* w, float* losses,
                     int batch, int64_t per_sample, csd_stream_t stream);

/* Inpainting projection after each predictor / corrector update (get_pc_inpainter, sampling/unconditional.py:266-277):
 * masked = mean_coef[b]*data + std_dev[b]*z; x_out = x*(1-mask) + masked*mask; x_mean = x_out*(1-mask) +
 * mean_coef[b]*data*mask. mask has the shape of data. x_out may alias x. mean_coef NULL = 1 (VE SDEs).            */
int csd_inpaint_merge_f32(const float* x, const float* data, const float* z, const float* mask, float* x_out, float* x_mean,
                          int batch, int64_t per_sample, const float* mean_coef, const float* std_dev, csd_stream_t stream);

/* dst[i] = table value for sample i at the current step (time labels, 1/sigma row scales). */
int csd_broadcast_table_f32(float* dst, int n, const float* tab, const int* step_idx, int sample_stride,
                            csd_stream_t stream);

/* *step_idx += 1 (one thread). */
int csd_step_advance(int* step_idx, csd_stream_t stream);

/* ---- score-network building blocks (NHWC bf16) ----------------------------------------------- */

/* out[b,h,w,c] (bf16, pitch c_pad, zero padded) = scale*src[b,c,h,w] + shift, channels taken from up
 * to two NCHW fp32 sources (x then y: NCSNpp_paired.forward cat, models/ncsnpp.py:395-398;
 * scale=2, shift=-1 is the `2*x-1` of models/ncsnpp.py:264-266).                               */
int csd_nchw_to_nhwc_bf16(const float* src0, int c0, const float* src1, int c1, void* out, int c_pad,
                          int batch, int h, int w, float scale, float shift, csd_stream_t stream);

/* dst[b,c,h,w] fp32 = src[b,h,w,c_off+c] * (row_scale ? row_scale[b] : 1) for c < c_cnt.
 * row_scale[b] carries the 1/sigma(t_b) of divide_by_sigmas (models/utils.py:50-74).           */
int csd_nhwc_bf16_to_nchw(const void* src, int c_pitch, int c_off, int c_cnt, float* dst, int batch,
                          int h, int w, const float* row_scale, csd_stream_t stream);

/* Per-channel GroupNorm statistics of ONE NHWC bf16 tensor: chan_sums[b, c, 0..1] = (sum x, sum x^2), stored (not
 * accumulated) by a fixed-order reduction: bitwise reproducible run to run. A tensor's sums are computed once and
 * reused by every GroupNorm that reads it (nn.GroupNorm call sites: models/layerspp.py:67,219,231; ncsnpp.py:200-233). */
int csd_gn_chan_stats_bf16(const void* src, int c, int pitch, float* chan_sums, int batch, int hw,
                           csd_stream_t stream);

/* chan_sums[b, c, :] = sum over the image's pixel tiles of the per-tile partials that csd_conv_gemm
 * (mode 2) writes from its epilogue: partials [batch * tiles_per_img, c, 2]. No atomics.          */
int csd_gn_finalize_partials_f32(const float* partials, float* chan_sums, int batch, int tiles_per_img, int c,
                                 csd_stream_t stream);

/* out = [SiLU](GroupNorm(cat(src0, src1))) as bf16 NHWC (pitch out_pitch >= c0 + c1, multiple of 8):
 * the `torch.cat([h, hs.pop()], dim=1)` + GroupNorm + SiLU of models/ncsnpp.py:325 / layerspp.py:242
 * without materialising the concatenation first. sums0 / sums1 are the per-channel sums of each source;
 * gamma / beta fp32 [c0 + c1]; eps as in the module (1e-6).                                       */
int csd_gn_apply_bf16(const void* src0, int c0, int pitch0, const float* sums0, const void* src1, int c1,
                      int pitch1, const float* sums1, const float* gamma, const float* beta, void* out,
                      int out_pitch, int batch, int hw, int groups, float eps, int apply_silu,
                      csd_stream_t stream);

/* One-launch form of csd_gn_chan_stats_bf16 + csd_gn_apply_bf16 for small images (the <= 20 px levels of
 * NCSN++ / the DDPM U-Net, where no producer delivers channel sums for free and three latency-bound launches
 * per nn.GroupNorm (models/layerspp.py:67,219,231,242) cost more than the data movement): one CTA per
 * (image, slice of whole groups) reads its sub-tensor once into registers, reduces the statistics on chip and
 * writes [SiLU](GroupNorm(cat(src0, src1))). csd_gn_fused_supported is pure host logic (1 = the shape fits:
 * hw * slice width within the register budget); csd_gn_fused_bf16 returns CSD_ERR_UNSUPPORTED otherwise.   */
int csd_gn_fused_supported(int c0, int c1, int hw, int groups, int batch);
int csd_gn_fused_bf16(const void* src0, int c0, int pitch0, const void* src1, int c1, int pitch1,
                      const float* gamma, const float* beta, void* out, int out_pitch, int batch, int hw,
                      int groups, float eps, int apply_silu, csd_stream_t stream);

/* Fused-prologue form of the GroupNorm above: instead of writing the normalised tensor, write per (image,
 * channel) the pair (scale, shift) = (rstd_g * gamma_c, beta_c - mean_g * rstd_g * gamma_c) of the GroupNorm
 * over cat(src0, src1), split per source: coef0 [batch, c0, 2], coef1 [batch, c1, 2] fp32. A csd_conv_gemm
 * segment that carries such a table (csd_conv_segment.norm) applies y = SiLU(x * scale + shift) to its
 * operand tile in shared memory, so GroupNorm -> SiLU -> conv3x3 (models/layerspp.py:242-266) is one kernel
 * and the normalised activation never exists in HBM.                                              */
int csd_gn_coeffs_f32(const float* sums0, int c0, const float* sums1, int c1, const float* gamma, const float* beta,
                      float* coef0, float* coef1, int batch, int hw, int groups, float eps, csd_stream_t stream);

/* csd_gn_coeffs_f32 with csd_gn_finalize_partials_f32 folded in: each source arrives either as channel sums
 * (sumsN) or as the per-tile partials of the transposed convolution's epilogue (partialsN [batch * tilesN, cN, 2],
 * exactly one of the two non-null); reduced partials are also written to sums_outN (may be null) for the tensor's
 * later consumers. One launch per fused GroupNorm instead of two.                                   */
int csd_gn_coeffs_partials_f32(const float* sums0, const float* partials0, int tiles0, float* sums_out0, int c0,
                               const float* sums1, const float* partials1, int tiles1, float* sums_out1, int c1,
                               const float* gamma, const float* beta, float* coef0, float* coef1, int batch, int hw,
                               int groups, float eps, csd_stream_t stream);

/* Depthwise separable FIR resampling of an NHWC bf16 tensor with the [1,3,3,1] family
 * (up_or_down_sampling.upsample_2d / downsample_2d, models/up_or_down_sampling.py:195-257):
 * mode 1 = up x2 (pad (2,1), gain 4), mode 2 = down x2 (pad (1,1)), mode 3 = same-rate pre-filter with
 * pad (2,2) (output (h+1) x (w+1); first half of conv_downsample_2d, :144-178). taps: 4 fp32 host values
 * (the un-normalised 1-D filter). If add != NULL, out = fir(src) + add (output-skip pyramid,
 * models/ncsnpp.py:344-349).                                                                    */
int csd_fir_resample_nhwc_bf16(const void* src, void* out, const void* add, int batch, int h, int w,
                               int c_pitch, int mode, const float* taps4_host, csd_stream_t stream);
/* The same with GroupNorm(+SiLU) applied to the INPUT on the fly: out = FIR(act(src * scale + shift)) [+ add], with the
 * per-(image, channel) fp32 (scale, shift) table `norm` [batch, norm_c, 2] of csd_gn_coeffs_* (norm_c a multiple of 8,
 * channels >= norm_c are resampled as stored). This is act(GroupNorm_0(x)) followed by upsample_2d / downsample_2d in
 * ResnetBlockBigGANpp (models/layerspp.py:242-258) without the normalised tensor in HBM; pixels outside the image
 * contribute zero, as the reference's zero-padded FIR of the activated tensor does. Modes 1 and 2 with a 16-byte aligned
 * source only (the TMA-staged kernel): CSD_ERR_UNSUPPORTED otherwise. *_f32: the fp32-activation plan.               */
int csd_fir_norm_resample_nhwc_bf16(const void* src, void* out, const void* add, const float* norm, int norm_c,
                                    int norm_silu, int batch, int h, int w, int c_pitch, int mode,
                                    const float* taps4_host, csd_stream_t stream);
int csd_fir_norm_resample_nhwc_f32(const void* src, void* out, const void* add, const float* norm, int norm_c,
                                   int norm_silu, int batch, int h, int w, int c_pitch, int mode,
                                   const float* taps4_host, csd_stream_t stream);

/* ---- fp32-activation variants ("tf32" plan: NHWC fp32 tensors, channel pitch a multiple of 8) ----
 * Same arguments and semantics as their *_bf16 namesakes above; src / out / add / probs are float. tcgen05 kind::tf32
 * truncates the fp32 words it reads, so producers of tensors that ONLY feed tensor-core operands round them to tf32
 * (nearest) on request: apply_silu bit 1 (csd_gn_apply_f32 / csd_gn_fused_f32), mode | 0x10 (csd_fir_resample_nhwc_f32),
 * always in csd_nchw_to_nhwc_f32 (the network input) and csd_softmax_rows_f32_f32 (attention probabilities). Together with
 * csd_conv_gemm's dtype = 1 they form the reference-precision plan (fp32 storage like the reference,
 * sampling/unconditional.py:206, models/ncsnpp.py:264-266; TF32 tensor-core operands like its cuDNN convolutions). */
int csd_nchw_to_nhwc_f32(const float* src0, int c0, const float* src1, int c1, void* out, int c_pad, int batch, int h,
                         int w, float scale, float shift, csd_stream_t stream);
int csd_nhwc_f32_to_nchw(const void* src, int c_pitch, int c_off, int c_cnt, float* dst, int batch, int h, int w,
                         const float* row_scale, csd_stream_t stream);
int csd_gn_chan_stats_f32(const void* src, int c, int pitch, float* chan_sums, int batch, int hw, csd_stream_t stream);
int csd_gn_apply_f32(const void* src0, int c0, int pitch0, const float* sums0, const void* src1, int c1, int pitch1,
                     const float* sums1, const float* gamma, const float* beta, void* out, int out_pitch, int batch,
                     int hw, int groups, float eps, int apply_silu, csd_stream_t stream);
int csd_gn_fused_supported_f32(int c0, int c1, int hw, int groups, int batch);
int csd_gn_fused_f32(const void* src0, int c0, int pitch0, const void* src1, int c1, int pitch1, const float* gamma,
                     const float* beta, void* out, int out_pitch, int batch, int hw, int groups, float eps,
                     int apply_silu, csd_stream_t stream);
int csd_fir_resample_nhwc_f32(const void* src, void* out, const void* add, int batch, int h, int w, int c_pitch,
                              int mode, const float* taps4_host, csd_stream_t stream);
int csd_softmax_rows_f32_f32(const float* logits, int in_pitch, void* probs, int out_pitch, int64_t rows, int cols,
                             float scale, csd_stream_t stream);

/* rows x cols softmax of fp32 logits (pitch in_pitch) times `scale`, bf16 output (pitch out_pitch,
 * columns >= cols zero-filled): F.softmax of models/layerspp.py:82-85.                          */
int csd_softmax_rows_f32_bf16(const float* logits, int in_pitch, void* probs, int out_pitch, int64_t rows,
                              int cols, float scale, csd_stream_t stream);

/* Time embedding (models/ncsnpp.py:242-260, models/layers.py:524-538, layerspp.py:32-41):
 * embedding_type 0 = positional(labels, nf), 1 = Gaussian Fourier (fourier_w [nf]);
 * temb = W1 @ act(W0 @ emb + b0) + b1, then act_temb = SiLU(temb) [batch, 4nf] fp32.            */
int csd_time_embedding_f32(const float* labels, int batch, int nf, int embedding_type,
                           const float* fourier_w, const float* w0, const float* b0, const float* w1,
                           const float* b1, float* act_temb, csd_stream_t stream);

/* All per-block Dense_0 projections in one pass (models/layerspp.py:262-263):
 * out[b, j] = dot(act_temb[b,:], w[j,:]) + bias[j], w [total_out, in_dim] fp32.                 */
int csd_dense_rows_f32(const float* act_temb, const float* w, const float* bias, float* out, int batch,
                       int in_dim, int total_out, csd_stream_t stream);

/* ---- tcgen05 implicit-GEMM convolution / GEMM -------------------------------------------------
 * D[m, n] = sum over segments s, taps t, channels c of A_s[pixel(m)+offset(t), c] * Wt[n, k(s,t,c)]
 * with A_s NHWC bf16 tensors (zero padded outside the image by TMA), Wt [n_rows, k_total] bf16
 * K-major. 3x3 taps are offsets (dy,dx) in {-1,0,1}^2 in row-major (ky,kx) order (torch conv2d
 * cross-correlation); 1 tap = 1x1 conv / plain GEMM. Epilogue: out = (D + bias + temb[b] + res)*scale.
 * Replaces nn.Conv2d (models/layers.py:100-132), NIN (models/layers.py:555-564), the attention
 * einsums (models/layerspp.py:82-86) and the `h += Dense_0(...)`, skip and 1/sqrt(2) of
 * ResnetBlockBigGANpp.forward (models/layerspp.py:260-274).                                     */
#define CSD_MAX_SEGMENTS 4

typedef struct csd_conv_segment {
  const void* a;       /* NHWC bf16 tensor [batch, h, w, pitch]                         */
  int32_t pitch;       /* channel pitch of `a` in elements (multiple of 8)              */
  int32_t c_off;       /* first channel used                                            */
  int32_t c_cnt;       /* channels used (zero-filled up to the next multiple of 32)     */
  int32_t taps;        /* 1 or 9                                                         */
  const float* norm;   /* mode 2 only: [batch, c_cnt, 2] (scale, shift) from csd_gn_coeffs_f32; the kernel
                          feeds act(x * scale + shift) to the MMA instead of x. NULL = raw operand      */
  int32_t norm_silu;   /* 1 = SiLU after the affine map (GroupNorm -> SiLU), 0 = affine only           */
  int32_t reserved_;
} csd_conv_segment;

typedef struct csd_conv_gemm_desc {
  int32_t batch, h, w;          /* spatial extent of the OUTPUT (tiles cover output pixels) */
  int32_t in_h, in_w;           /* spatial extent of A (0 = same as h, w)                  */
  int32_t stride;               /* conv stride (1 or 2); taps read in[o*stride + k - pad]  */
  int32_t pad;                  /* 1 = 'same' 3x3 (torch padding=1), 0 = valid / padded-after
                                   (DDPM Downsample pads bottom/right: TMA zero-fills it)   */
  int32_t tile_w, tile_h, tile_b; /* pixel box per CTA, product <= 128 (ignored in halo mode) */
  int32_t mode;                 /* 0 = one TMA box per tap; 1 = halo: 16x8-pixel tiles, one halo load per
                                   32-channel chunk feeds all 9 taps (3x3, stride 1, pad 1 only);
                                   2 = transposed halo: M = 128 output channels, N = 32x8 pixels     */
  int32_t mt;                   /* halo mode: vertically stacked tiles per CTA (1..4), mt*n_tile <= 512 */
  int32_t nseg;
  int32_t n;                    /* output columns computed (rows of Wt used)             */
  int32_t n_store;              /* output columns written (>= n allowed: zero columns)   */
  int32_t n_tile;               /* columns per CTA: multiple of 16, <= 512                */
  csd_conv_segment seg[CSD_MAX_SEGMENTS];
  const void* wt;               /* bf16 [wt_batch?][wt_rows, k_total]                     */
  int32_t wt_rows;              /* rows available (>= ceil(n / n_tile) * n_tile)         */
  int32_t k_total;              /* sum over segments of taps * ceil32(c_cnt)             */
  int32_t wt_pitch;             /* row pitch of Wt in elements (0 = k_total)              */
  int32_t wt_k_off;             /* first K column of Wt used (attention: K inside [q|k])  */
  int32_t k_valid;              /* K columns of Wt that exist (0 = k_total); rest zero-filled */
  int64_t wt_batch_stride;      /* elements between z-batches of Wt (0 = shared weights)  */
  int32_t z_batches;            /* gridDim.z: batched GEMM count (1 for convolutions)    */
  int32_t a_batch_step;         /* A batch coordinate += z * a_batch_step                 */
  void* out;                    /* bf16 or fp32, addressed [pixel * out_pitch + n]        */
  int32_t out_pitch;
  int32_t out_f32;              /* 0 = bf16, 1 = fp32                                     */
  int64_t out_z_stride;         /* elements between z-batches of out                      */
  const float* bias;            /* fp32, padded to the n_tile multiple, or NULL           */
  int32_t bias_per_row;         /* 0 = bias[n], 1 = bias[m] (row index inside the z batch) */
  const float* temb;            /* fp32 [batch, temb_pitch] (already offset), or NULL      */
  int32_t temb_pitch;
  const void* res;              /* bf16 residual addressed like out (pitch res_pitch), or NULL */
  int32_t res_pitch;
  int64_t res_z_stride;
  float scale;
  float* stat_partials;         /* mode 2 only: [pixel tiles, n_store, 2] per-tile per-channel (sum, sum of
                                   squares) of the stored bf16 output (GroupNorm statistics), or NULL */
  int32_t dtype;                /* 0 = bf16 activations / weights / residual, tcgen05 kind::f16 (the fast plan);
                                   1 = fp32 activations / weights / residual / output, tcgen05 kind::tf32 - the
                                   reference's own precision class (fp32 storage, sampling/unconditional.py:206; cuDNN
                                   TF32 convolutions): A, wt, res, out are float, pitches stay in elements      */
  int32_t out_round_tf32;       /* dtype 1: round the stored output to tf32 (nearest): set for tensors that only feed
                                   further tensor-core operands (attention q|k, V^T, P V), whose fp32 words the
                                   kind::tf32 MMA would otherwise truncate                                      */
  int32_t k_splits;             /* mode 0, z_batches 1: > 1 splits the K range (taps x channel chunks) over gridDim.z
                                   CTAs per tile - for levels with fewer tiles than SMs. Partial sums go to splitk_ws
                                   and a second launch adds them in split order and applies bias / temb / res / scale */
  float* splitk_ws;             /* k_splits > 1: fp32 workspace, k_splits * batch*h*w * ceil8(n_store) floats           */
} csd_conv_gemm_desc;

int csd_conv_gemm(const csd_conv_gemm_desc* desc, csd_stream_t stream);

/* Few-channel 3x3 output heads (conv3x3(SiLU(GroupNorm(h))) -> 3 / 6 channels, models/ncsnpp.py:337-352, 372-381) as a
 * tap-stacked 1x1 convolution + this pass: partial [batch, h, w, p_pitch] bf16 holds, at channel t * cout + co, tap t's
 * contribution computed at the UNSHIFTED pixel; out[b, y, x, co] = bias[co] + res[b, y, x, co] + sum over the 9 taps of
 * partial[b, y + t/3 - 1, x + t%3 - 1, t * cout + co] (zero outside the image). cout <= 8; out channels >= cout are
 * written as zero.                                                                                               */
int csd_tap_shift_sum_bf16(const void* partial, int p_pitch, int cout, const float* bias, const void* res, int res_pitch,
                           void* out, int out_pitch, int batch, int h, int w, csd_stream_t stream);

/* ---- fused self-attention core -----------------------------------------------------------------------
 * AttnBlockpp.forward after the q|k|v projection (models/layerspp.py:82-91) and the DDPM AttnBlock
 * (models/layers.py:583-590) as ONE tcgen05 kernel, one CTA per (image, 128-query tile):
 *   S = Q K^T (fp32 in TMEM) -> softmax over the L keys with logits * C^-0.5 -> O = P V -> Y = O Wo^T + bo ->
 *   out = (res + Y) * out_scale.
 * The [batch, L, L] logits / probabilities and the attention output never reach HBM.
 * qkv: bf16 [batch, L, qkv_pitch] with q | k | v at channels [0,C) | [C,2C) | [2C,3C) (written by one 1x1
 * csd_conv_gemm over the GroupNorm output); wo: bf16 K-major [wo_rows >= C, wo_pitch >= C] = NIN_3.W^T as packed
 * for csd_conv_gemm; bo fp32 [>= C]; res / out: bf16 [batch, L, pitch] (x and the block output; out may alias
 * neither input). L <= 512, C a multiple of 16 <= 320; csd_attn_core_supported tells whether (L, C) fits the kernel's
 * shared-memory plan - callers fall back to separate GEMM / softmax launches otherwise.                       */
int csd_attn_core_supported(int L, int C);
int csd_attn_core_bf16(const void* qkv, int qkv_pitch, const void* wo, int wo_pitch, int wo_rows, const float* bo,
                       const void* res, int res_pitch, void* out, int out_pitch, int batch, int L, int C,
                       float out_scale, csd_stream_t stream);

/* ---- device-resident adaptive Runge-Kutta pieces ---------------------------------------------
 * The reference integrates the probability-flow ODE with scipy's RK45 on the host: the full state crosses PCIe twice
 * per right-hand side (likelihood.py:91-99, sampling/unconditional.py:140-150). With these two kernels the state and
 * the stage derivatives stay in HBM; only the 4-byte error norm returns to the host for step-size control.
 * k_stack is [stages, n] fp32; coefficient vectors are HOST arrays of `stages` floats (stages <= 8).            */
/* out = y + h * sum_s coef[s] * k_stack[s]                                                                     */
int csd_rk_combine_f32(const float* y, const float* k_stack, int64_t n, int stages, const float* coef_host, float h,
                       float* out, csd_stream_t stream);
/* Device-wide sums below are deterministic (fixed summation order, no float atomics): `out` is a workspace of
 * CSD_REDUCE_WS_FLOATS floats that the caller zeroes ONCE when it allocates it; out[0] receives the result, out[1] is
 * a ticket counter the kernel leaves at zero, out[8..] hold per-CTA partials.                                    */
#define CSD_REDUCE_WS_FLOATS 1024
/* out[0] = sum_i (h * sum_s e[s] k_stack[s][i] / (atol + max(|y_i|, |y2_i|) * rtol))^2 (device scalar)          */
int csd_rk_error_sumsq_f32(const float* k_stack, int64_t n, int stages, const float* e_host, float h, const float* y,
                           const float* y2, float atol, float rtol, float* out, csd_stream_t stream);

/* ---- fused optimizer step ----------------------------------------------------------------
 * The reference runs torch.nn.utils.clip_grad_norm_ + optim.Adam.step (losses.py:38-52) and then
 * ExponentialMovingAverage.update (models/ema.py:64-93: a Python loop of 3 kernels per parameter tensor) after every
 * training step. Over a flat fp32 parameter buffer this is one reduction and one elementwise pass.              */
/* out[0] = sum x^2 (device scalar; `out` = CSD_REDUCE_WS_FLOATS workspace, see csd_rk_error_sumsq_f32).        */
int csd_sumsq_f32(const float* x, int64_t n, float* out, csd_stream_t stream);
/* g' = g * min(1, max_norm / (sqrt(*gnorm_sq) + 1e-6)) (max_norm < 0: no clipping) + weight_decay * p;
 * m = beta1 m + (1-beta1) g'; v = beta2 v + (1-beta2) g'^2; p -= lr / bias_corr1 * m / (sqrt(v / bias_corr2) + eps);
 * ema -= (1 - ema_decay) * (ema - p) (ema may be NULL). torch.optim.Adam (no amsgrad) semantics.                */
int csd_fused_adam_ema_f32(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr, float beta1,
                           float beta2, float eps, float weight_decay, float bias_corr1, float bias_corr2, float max_norm,
                           const float* gnorm_sq, float ema_decay, csd_stream_t stream);

/* ---- batched weight packing -------------------------------------------------------------------------------
 * One job = one block of a packed operand. kind 0: dst (bf16) [(r)*dst_pitch + tap*k_pad + c] = scale * src[r*s_row +
 * c*s_col + ts*s_tap], ts = flip ? taps-1-tap : tap, for r < rows, c < cols, tap < taps (dst / src already offset to
 * the block; padding elements of dst are never written and stay zero). Forward nn.Conv2d weight [Cout,Cin,kh,kw]: r =
 * co, c = ci, (s_row, s_col, s_tap) = (Cin*taps, taps, 1); its data gradient: r = ci, c = co, flip = 1; NIN.W [in,out]:
 * (1, out, 0). kind 1: dst (fp32) [i] = src[i] + src2[i] for i < rows (bias vectors; src2 may be NULL). kind 2: kind 0
 * with an fp32 destination (operands of the tf32 plan, csd_conv_gemm dtype = 1).                                   */
typedef struct csd_pack_job {
  const float* src;
  const float* src2;
  void* dst;
  int64_t s_row, s_col, s_tap;
  int32_t rows, cols, taps, k_pad, dst_pitch, flip, kind;
  float scale;
} csd_pack_job;

/* Runs all jobs of a device-resident table in one launch (max_elems = largest rows*cols*taps, sizes the grid). */
int csd_pack_weights(const csd_pack_job* jobs_dev, int njobs, int64_t max_elems, csd_stream_t stream);

/* ---- training backward ------------------------------------------------------------------------------
 * The reference obtains every gradient from PyTorch autograd (`loss.backward()` under Lightning,
 * lightning_modules/BaseSdeGenerativeModel.py:57-60; losses.py:345-407 in the non-Lightning step_fn): cuDNN
 * dgrad / wgrad for nn.Conv2d, ATen backward for nn.GroupNorm, nn.SiLU, F.softmax, einsum, and the custom
 * autograd.Function of upfirdn2d (op/upfirdn2d.py:19-142). The entry points below are the adjoints of this
 * library's forward kernels. Activation gradients are NHWC bf16 (same pitch as the activation), parameter
 * gradients fp32 in the reference's parameter layout. Data gradients of convolutions (dgrad) are csd_conv_gemm
 * calls with flipped / transposed packed weights (stride-2 convolutions through csd_zero_stuff_nhwc_bf16).   */

/* Pixel-major operand layout of the weight-gradient GEMM: one row per channel, pixels contiguous on a zero
 * padded (grid_h + 2) x wp grid, `ips` images per split-K slice; rows are [margin | kp | margin] elements.     */
typedef struct csd_pixmajor_geom {
  int64_t q;          /* padded grid pixels per image = (grid_h + 2) * wp                */
  int64_t kp;         /* K extent of one split (ips * q rounded up to 64)                */
  int64_t row_pitch;  /* elements per channel row = margin + kp + margin                  */
  int32_t wp;         /* padded row width = ceil8(grid_w + 2)                            */
  int32_t ips;        /* images per split                                                */
  int32_t splits;     /* ceil(batch / ips)                                               */
  int32_t margin;     /* zero columns before / after the data (>= wp + 8, multiple of 64) */
} csd_pixmajor_geom;

int csd_pixmajor_geometry(int batch, int grid_h, int grid_w, csd_pixmajor_geom* g);

/* out[copy][split][c][margin + bl*q + (y*stride+offset+1)*wp + (x*stride+offset) + 2 - copy_kx] = src[b, y, x, c_off+c]
 * with b = split*ips + bl. ncopies = 3 writes the kx = 0,1,2 shifted copies a 3x3 contraction needs (copy index =
 * kx), ncopies = 1 only the centre (kx = 1). `out` must have been zero-filled once: positions the pattern
 * does not touch are the grid's zero padding and are never written.                                          */
int csd_nhwc_to_pixmajor_bf16(const void* src, int pitch, int c_off, int c_cnt, int batch, int h, int w, int stride,
                              int offset, const csd_pixmajor_geom* g, int ncopies, void* out, csd_stream_t stream);

/* partial[split][tap][co][ci] = sum over the split's grid pixels q of g_pm[split][co][q] * a_pm[kx][split][ci][q +
 * (ky-1)*wp] (tcgen05, fp32 accumulation): the wgrad of nn.Conv2d 3x3 (taps = 9) / 1x1 and NIN (taps = 1).     */
int csd_wgrad_gemm_bf16(const void* g_pm, int cout, const void* a_pm, int cin, int taps, const csd_pixmajor_geom* g,
                        float* partial, csd_stream_t stream);

/* The same partial sums straight from the NHWC tensors (no pixel-major copies): g [batch, h, w, g_pitch] is the output
 * gradient ON THE ACTIVATION GRID (the zero-stuffed gradient for stride-2 convolutions), a [batch, h, w, a_pitch] the
 * activation; channels [g_c_off, +cout) x [a_c_off, +cin). MN-major tcgen05 operands, one TMA halo load feeds the three
 * kx taps of a CTA. Needs >= 64-channel pitches (narrower tensors use the pixel-major path). `splits` must be the value
 * csd_wgrad_direct_splits returns for the same shape: partial is [splits][taps][cout][cin].                      */
int csd_wgrad_direct_splits(int batch, int h, int w, int cout, int cin, int taps, int* splits);
int csd_wgrad_direct_bf16(const void* g, int g_pitch, int g_c_off, int cout, const void* a, int a_pitch, int a_c_off,
                          int cin, int taps, int batch, int h, int w, float* partial, int splits, csd_stream_t stream);

/* dw[co*stride_co + (ci_off+ci)*stride_ci + tap*stride_tap] (+)= scale * sum_split partial[split][tap][co][ci]
 * (fixed summation order). nn.Conv2d weight [Cout, Cin, 3, 3]: strides (Cin*9, 9, 1); NIN.W [in, out]: (1, out, 0). */
int csd_wgrad_reduce_f32(const float* partial, int splits, int taps, int cout, int cin, float scale, float* dw,
                         int64_t stride_co, int64_t stride_ci, int64_t stride_tap, int ci_off, int accumulate,
                         csd_stream_t stream);

/* GroupNorm(+SiLU) backward in three passes (adjoint of csd_gn_apply_bf16 / the fused conv prologue):
 * stats:  s[b, s_c_off + c, 0..1] += (sum_p du, sum_p du * x), du = dy * silu'(x*scale + shift) (or dy), per source;
 *         fwd_coef = that source's (scale, shift) table from csd_gn_coeffs_f32; the caller zeroes s beforehand.
 * coeffs: bwd_coef[b, c, 0..3] = (A, B, C, rstd) with dx = A*du + B*x + C over the channel concatenation; if dgamma /
 *         dbeta are given, dgamma[c] += sum_b rstd*(S2 - mean*S1), dbeta[c] += sum_b S1 (s is overwritten with the
 *         per-image contributions on the way).
 * apply:  dx (=|+=) A*du + B*x + C for one source.                                                            */
int csd_gn_bwd_stats_bf16(const void* x, int c, int x_pitch, const void* dy, int dy_pitch, int dy_c_off,
                          const float* fwd_coef, float* s, int c_total, int s_c_off, int batch, int hw, int silu,
                          csd_stream_t stream);
int csd_gn_bwd_coeffs_f32(const float* sums0, int c0, const float* sums1, int c1, const float* gamma, float* s,
                          float* bwd_coef, float* dgamma, float* dbeta, int batch, int hw, int groups, float eps,
                          csd_stream_t stream);
int csd_gn_bwd_apply_bf16(const void* x, int c, int x_pitch, const void* dy, int dy_pitch, int dy_c_off,
                          const float* fwd_coef, const float* bwd_coef, int c_total, int b_c_off, void* dx, int dx_pitch,
                          int batch, int hw, int silu, int accumulate, csd_stream_t stream);

/* Adjoint of csd_fir_resample_nhwc_bf16 (UpFirDn2dBackward, op/upfirdn2d.py:19-85): g is the gradient of the
 * forward OUTPUT, din [batch, h, w, c_pitch] the gradient of the forward input (h, w = forward input extent),
 * same mode numbers and taps as the forward call.                                                            */
int csd_fir_resample_bwd_nhwc_bf16(const void* g, void* din, int batch, int h, int w, int c_pitch, int mode,
                                   const float* taps4_host, int accumulate, csd_stream_t stream);

/* ds = scale * p * (dp - sum_j p_j dp_j) per row: adjoint of csd_softmax_rows_f32_bf16 (dp fp32, ds bf16 with
 * columns >= cols zero-filled up to ds_pitch).                                                               */
int csd_softmax_bwd_bf16(const void* probs, int p_pitch, const float* dp, int dp_pitch, void* ds, int ds_pitch,
                         int64_t rows, int cols, float scale, csd_stream_t stream);

/* out[z, c, r] = in[z, r, c] (bf16; the K-major operands of the attention backward GEMMs).                   */
int csd_transpose_bf16(const void* in, int in_pitch, int64_t in_z_stride, void* out, int out_pitch, int64_t out_z_stride,
                       int rows, int cols, int z, csd_stream_t stream);

/* dst = alpha * src (+ dst if accumulate): gradient of a residual / skip addition. n elements, multiple of 8. */
int csd_axpy_bf16(const void* src, void* dst, int64_t n, float alpha, int accumulate, csd_stream_t stream);

/* nn.Dropout of the ResNet blocks (models/layerspp.py:266, models/layers.py:664) in train mode: out[i] = keep_i ?
 * x[i] / (1 - p) : 0 with keep_i a counter-based hash of (*seed_dev, salt, i). Calling it on the gradient with the
 * same (seed, salt) is the backward pass; no mask is stored. out may alias x. n multiple of 8.                 */
int csd_dropout_bf16(const void* x, void* out, int64_t n, float p, const int64_t* seed_dev, uint64_t salt,
                     csd_stream_t stream);

/* dst[b, y*stride+offset, x*stride+offset, :] = src[b, y, x, :], zero elsewhere: turns the data gradient of a
 * stride-2 convolution into a stride-1 convolution with the flipped weights.                                  */
int csd_zero_stuff_nhwc_bf16(const void* src, void* dst, int batch, int h, int w, int dst_h, int dst_w, int c_pitch,
                             int stride, int offset, csd_stream_t stream);

/* Adjoint of csd_nhwc_bf16_to_nchw: out[b,h,w,c] = g0[b,c,h,w]*row_scale0[b] (c < c0), g1[...]*row_scale1[b]
 * (c0 <= c < c0+c1), 0 for padding channels. g0 / g1 may be NULL (no gradient for that output group).          */
int csd_nchw_grad_to_nhwc_bf16(const float* g0, int c0, const float* row_scale0, const float* g1, int c1,
                               const float* row_scale1, void* out, int c_pad, int batch, int h, int w,
                               csd_stream_t stream);

/* From chan_sums [batch, sums_c, 2] of an output gradient (csd_gn_chan_stats_bf16; sums_c >= c, e.g. the channel
 * pitch): dbias0/1[ch] += scale * sum_b sums[b, ch], dtproj[b*tproj_pitch + ch] += scale * sums[b, ch] for ch < c
 * (gradient of the bias and of the `h += Dense_0(act(temb))` add). Any output may be NULL.                      */
int csd_bias_temb_grad_f32(const float* chan_sums, int sums_c, int batch, int c, float scale, float* dbias0,
                           float* dbias1, float* dtproj, int tproj_pitch, csd_stream_t stream);

/* Small fp32 GEMM for the time-embedding MLP / Dense_0 projections and their gradients:
 * C = alpha * op(A) op(B) + beta * C + bias[n], op(A)(m,k) = trans_a ? A[k*lda+m] : A[m*lda+k], op(B)(k,n) =
 * trans_b ? B[n*ldb+k] : B[k*ldb+n].                                                                          */
int csd_sgemm_small_f32(int trans_a, int trans_b, int m, int n, int k, float alpha, const float* a, int lda,
                        const float* b, int ldb, float beta, float* c, int ldc, const float* bias, csd_stream_t stream);

/* y = silu(x) (grad = 0) or y = dy * silu'(x) (grad = 1), fp32.                                               */
int csd_silu_f32(const float* x, const float* dy, float* y, int64_t n, int grad, csd_stream_t stream);

/* The sinusoidal / Gaussian-Fourier features alone (first stage of csd_time_embedding_f32): emb [batch, nf] or
 * [batch, 2nf].                                                                                              */
int csd_time_features_f32(const float* labels, int batch, int nf, int embedding_type, const float* fourier_w, float* emb,
                          csd_stream_t stream);

/* Adjoint of csd_dsm_loss_f32: dscore[b,i] = grad_losses[b] * 2*w[b]*a[b] * (a[b]*score[b,i] + c[b]*z[b,i]).   */
int csd_dsm_loss_bwd_f32(const float* score, const float* z, const float* a, const float* c, const float* w,
                         const float* grad_losses, float* dscore, int batch, int64_t per_sample, csd_stream_t stream);


#ifdef __cplusplus
}
#endif

#endif /* CSD_B200_H_ */
