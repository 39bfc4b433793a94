// upfirdn2d (NCHW fp32 planes, the reference's native op) and the NHWC bf16 FIR resampler used
// inside the score network.
//
// Reference: op/upfirdn2d_kernel.cu:49-207 (kernels), :209-369 (dispatch), op/upfirdn2d.py:159-200
// (semantics). Both reference kernels use scalar 4-byte loads, `volatile` shared memory and one
// block shape for every size. Here:
//   * tile kernel  - one CTA = one output tile of one plane. The input footprint of the tile is
//                    staged in shared memory by ONE TMA box load (cp.async.bulk.tensor.3d over
//                    [planes, H, W]); coordinates outside the image are zero-filled by the TMA unit,
//                    which is exactly upfirdn's zero padding. Outputs are evaluated polyphase (only
//                    the K/up taps that hit a real sample) and written as 8/16-byte vectors.
//                    Instantiated for the three geometries NCSN++ uses (up2 / down2 / 1:1, 4x4).
//   * generic kernel - any up/down/pad/kernel <= 8x8, one thread per output. Also the fallback when
//                    a row pitch is not a multiple of 16 bytes (TMA's stride rule).
// HBM-bound: algorithmic bytes = 4 * (in + out elements).
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"

#include <type_traits>
#include "tensormap.cuh"
#include "../../include/csd_b200.h"

namespace csd {

struct UpfirdnParams {
  const float* in;
  const float* kernel;  // device [kh, kw]
  float* out;
  long long planes;
  int in_h, in_w, out_h, out_w;
  int kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_y0;
  int tiles_x, tiles_y;
};

__device__ __forceinline__ int pmod(int a, int m) {
  int r = a % m;
  return r < 0 ? r + m : r;
}
__device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// ---- generic: one thread per output ------------------------------------------------------------
__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(UpfirdnParams p) {
  __shared__ float kf[64];
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    const int ky = i / p.kw, kx = i % p.kw;
    kf[i] = p.kernel[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];  // flipped: true convolution
  }
  __syncthreads();
  const long long total = p.planes * p.out_h * p.out_w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % p.out_w);
    const int oy = (int)((idx / p.out_w) % p.out_h);
    const long long plane = idx / ((long long)p.out_w * p.out_h);
    const float* src = p.in + plane * p.in_h * p.in_w;
    const int uy = oy * p.down_y - p.pad_y0;
    const int ux = ox * p.down_x - p.pad_x0;
    float acc = 0.f;
    for (int ky = pmod(-uy, p.up_y); ky < p.kh; ky += p.up_y) {
      const int iy = (uy + ky) / p.up_y;  // exact: uy + ky is a multiple of up_y
      if (uy + ky < 0 || iy >= p.in_h) continue;
      for (int kx = pmod(-ux, p.up_x); kx < p.kw; kx += p.up_x) {
        const int ix = (ux + kx) / p.up_x;
        if (ux + kx < 0 || ix >= p.in_w) continue;
        acc += kf[ky * p.kw + kx] * __ldg(src + (long long)iy * p.in_w + ix);
      }
    }
    p.out[idx] = acc;
  }
}

// ---- tiled: TMA-staged input footprint, polyphase evaluation, vector stores ----------------------
template <int UP, int DOWN, int K>
struct TileCfg {
  static constexpr int TX = 32, TY = 8;
  static constexpr int CX = (DOWN == 1) ? 4 : 2;  // outputs per thread along x (one vector store)
  static constexpr int RY = 4;                    // output rows per thread
  static constexpr int TOW = TX * CX, TOH = TY * RY;
  static constexpr int SPAN_W = (TOW - 1) * DOWN + K, SPAN_H = (TOH - 1) * DOWN + K;
  // +3: the tile starts at a column that is a multiple of 4 (TMA needs 16-byte aligned box rows)
  static constexpr int IN_TW = (((SPAN_W + UP - 1) / UP + 1) + 3 + 3) & ~3;
  static constexpr int IN_TH = (SPAN_H + UP - 1) / UP + 1;
  static_assert(IN_TW <= 256 && IN_TH <= 256, "TMA box dimension limit");
};

template <int UP, int DOWN, int K, bool USE_TMA>
__global__ void __launch_bounds__(256)
upfirdn2d_tile_kernel(const __grid_constant__ CUtensorMap map, const UpfirdnParams p) {
  using C = TileCfg<UP, DOWN, K>;
  __shared__ __align__(128) float tile[C::IN_TH * C::IN_TW];
  __shared__ float kf[K * K];
  __shared__ __align__(8) unsigned long long bar;

  const int tid = threadIdx.x;
  long long bid = blockIdx.x;
  const int tix = (int)(bid % p.tiles_x);
  const int tiy = (int)((bid / p.tiles_x) % p.tiles_y);
  const long long plane = bid / ((long long)p.tiles_x * p.tiles_y);
  const int ox0 = tix * C::TOW, oy0 = tiy * C::TOH;
  // first input sample whose upsampled position is >= the first tap of the tile
  const int ix0 = floor_div(floor_div(ox0 * DOWN - p.pad_x0 + UP - 1, UP), 4) * 4;
  const int iy0 = floor_div(oy0 * DOWN - p.pad_y0 + UP - 1, UP);

  if (USE_TMA) {
    if (tid == 0) {
      ptx::mbar_init(ptx::smem_u32(&bar), 1);
      ptx::fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar), C::IN_TH * C::IN_TW * 4);
      ptx::tma_load_3d(ptx::smem_u32(tile), &map, ptx::smem_u32(&bar), ix0, iy0, (int)plane);
    }
  } else {
    const float* src = p.in + plane * p.in_h * p.in_w;
    for (int i = tid; i < C::IN_TH * C::IN_TW; i += 256) {
      const int r = i / C::IN_TW, c = i % C::IN_TW;
      const int gy = iy0 + r, gx = ix0 + c;
      tile[i] = (gy >= 0 && gy < p.in_h && gx >= 0 && gx < p.in_w) ? __ldg(src + (long long)gy * p.in_w + gx) : 0.f;
    }
  }
  if (tid < K * K) {
    const int ky = tid / K, kx = tid % K;
    kf[tid] = p.kernel[(K - 1 - ky) * K + (K - 1 - kx)];
  }
  if (USE_TMA) {
    __syncthreads();  // kf visible
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  } else {
    __syncthreads();
  }

  if constexpr (UP == 1) {
    // UP = 1 (down x2, 1:1): per-output evaluation straight from the shared-memory tile. The register-blocked form below
    // was measured SLOWER here (ncu r2: 100 vs 78 us for down x2): adjacent lanes sit 4 floats apart in the tile, so each
    // of its footprint loads is a 4-way bank conflict, and the footprint (60 / 49 samples) is not reused enough to pay.
    const int txx = tid & 31, tyy = tid >> 5;
    const int oxb = ox0 + txx * C::CX;
    constexpr int NT = (K + UP - 1) / UP;  // taps that can hit a real sample, per axis
    int kx0[C::CX], xi[C::CX];
  #pragma unroll
    for (int c = 0; c < C::CX; ++c) {
      const int ux = (oxb + c) * DOWN - p.pad_x0;
      kx0[c] = pmod(-ux, UP);
      xi[c] = (ux + kx0[c] - ix0 * UP) / UP;  // column in the tile of the first contributing sample
    }
    float* dst_plane = p.out + plane * p.out_h * p.out_w;
  #pragma unroll
    for (int r = 0; r < C::RY; ++r) {
      const int oy = oy0 + tyy * C::RY + r;
      if (oy >= p.out_h) break;
      const int uy = oy * DOWN - p.pad_y0;
      const int ky0 = pmod(-uy, UP);
      const int yi = (uy + ky0 - iy0 * UP) / UP;
      float acc[C::CX];
  #pragma unroll
      for (int c = 0; c < C::CX; ++c) acc[c] = 0.f;
  #pragma unroll
      for (int jy = 0; jy < NT; ++jy) {
        const int ky = ky0 + jy * UP;
        if (ky < K) {
          const float* srow = tile + (yi + jy) * C::IN_TW;
  #pragma unroll
          for (int c = 0; c < C::CX; ++c) {
  #pragma unroll
            for (int jx = 0; jx < NT; ++jx) {
              const int kx = kx0[c] + jx * UP;
              if (kx < K) acc[c] = fmaf(kf[ky * K + kx], srow[xi[c] + jx], acc[c]);
            }
          }
        }
      }
      float* dst = dst_plane + (long long)oy * p.out_w + oxb;
      const bool vec_ok = (oxb + C::CX <= p.out_w) && ((p.out_w & (C::CX - 1)) == 0);
      if (vec_ok) {
        if (C::CX == 4) {
          *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        } else {
          *reinterpret_cast<float2*>(dst) = make_float2(acc[0], acc[1]);
        }
      } else {
  #pragma unroll
        for (int c = 0; c < C::CX; ++c)
          if (oxb + c < p.out_w) dst[c] = acc[c];
      }
    }
  } else {
    // Register-blocked polyphase evaluation: a thread owns a CX x RY output block. Its input footprint (FW x FH samples of
    // the shared-memory tile) and the K x K filter are read into registers ONCE, then every output is a static sum over
    // them - 16 (up x2) / 60 (down x2) / 49 (1:1) shared-memory loads per block instead of 2 per multiply-add. For UP = 2
    // the tap phase of an output alternates with its coordinate; the phase of the block's first row / column (A, B) is
    // uniform over the warp, so the four phase combinations are four fully unrolled code paths.
    const int txx = tid & 31, tyy = tid >> 5;
    const int oxb = ox0 + txx * C::CX, oyb = oy0 + tyy * C::RY;
    const int ux0 = oxb * DOWN - p.pad_x0, uy0 = oyb * DOWN - p.pad_y0;
    const int B0 = pmod(-ux0, UP), A0 = pmod(-uy0, UP);
    const int xi0 = (ux0 + B0 - ix0 * UP) / UP, yi0 = (uy0 + A0 - iy0 * UP) / UP;
    constexpr int FW = UP == 1 ? (C::CX - 1) * DOWN + K : C::CX / 2 + 2;
    constexpr int FH = UP == 1 ? (C::RY - 1) * DOWN + K : C::RY / 2 + 2;
    static_assert(UP == 1 || (UP == 2 && DOWN == 1 && K == 4), "register-blocked evaluation: UP 1, or UP 2 with a 4-tap filter");
    float in[FH][FW];
  #pragma unroll
    for (int a = 0; a < FH; ++a)
  #pragma unroll
      for (int bb = 0; bb < FW; ++bb) in[a][bb] = tile[(yi0 + a) * C::IN_TW + xi0 + bb];
    float kr[K * K];
  #pragma unroll
    for (int i = 0; i < K * K; ++i) kr[i] = kf[i];
    float acc[C::RY][C::CX];
  #pragma unroll
    for (int r = 0; r < C::RY; ++r)
  #pragma unroll
      for (int c = 0; c < C::CX; ++c) acc[r][c] = 0.f;

    auto eval = [&](auto a_tag, auto b_tag) {
      constexpr int A = decltype(a_tag)::value, Bp = decltype(b_tag)::value;
  #pragma unroll
      for (int r = 0; r < C::RY; ++r) {
        // first tap row of output row r and its offset inside the footprint
        constexpr int dummy = 0;
        (void)dummy;
        const int ky0 = UP == 1 ? 0 : ((A + r) & 1);
        const int fy = UP == 1 ? r * DOWN : (r + ((A + r) & 1) - A) / 2;
  #pragma unroll
        for (int c = 0; c < C::CX; ++c) {
          const int kx0 = UP == 1 ? 0 : ((Bp + c) & 1);
          const int fx = UP == 1 ? c * DOWN : (c + ((Bp + c) & 1) - Bp) / 2;
          float s = 0.f;
  #pragma unroll
          for (int jy = 0; jy < (K + UP - 1) / UP; ++jy)
  #pragma unroll
            for (int jx = 0; jx < (K + UP - 1) / UP; ++jx)
              s = fmaf(kr[(ky0 + jy * UP) * K + kx0 + jx * UP], in[fy + jy][fx + jx], s);
          acc[r][c] = s;
        }
      }
    };
    if (UP == 1) {
      eval(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
    } else if (A0 == 0) {
      if (B0 == 0) eval(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
      else eval(std::integral_constant<int, 0>{}, std::integral_constant<int, 1>{});
    } else {
      if (B0 == 0) eval(std::integral_constant<int, 1>{}, std::integral_constant<int, 0>{});
      else eval(std::integral_constant<int, 1>{}, std::integral_constant<int, 1>{});
    }

    float* dst_plane = p.out + plane * p.out_h * p.out_w;
    const bool vec_ok = (oxb + C::CX <= p.out_w) && ((p.out_w & (C::CX - 1)) == 0);
  #pragma unroll
    for (int r = 0; r < C::RY; ++r) {
      const int oy = oyb + r;
      if (oy >= p.out_h) break;
      float* dst = dst_plane + (long long)oy * p.out_w + oxb;
      if (vec_ok) {
        if (C::CX == 4) {
          *reinterpret_cast<float4*>(dst) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        } else {
          *reinterpret_cast<float2*>(dst) = make_float2(acc[r][0], acc[r][1]);
        }
      } else {
  #pragma unroll
        for (int c = 0; c < C::CX; ++c)
          if (oxb + c < p.out_w) dst[c] = acc[r][c];
      }
    }
  }
}

template <int UP, int DOWN, int K>
static int launch_tile(const UpfirdnParams& p0, cudaStream_t stream) {
  using C = TileCfg<UP, DOWN, K>;
  UpfirdnParams p = p0;
  p.tiles_x = ceil_div(p.out_w, C::TOW);
  p.tiles_y = ceil_div(p.out_h, C::TOH);
  const long long blocks = p.planes * p.tiles_x * p.tiles_y;
  CSD_REQUIRE(blocks < (1LL << 31), "upfirdn2d: too many tiles (%lld)", blocks);
  const bool tma_ok = ((p.in_w * 4) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0) &&
                      p.planes < (1LL << 31) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (tma_ok) {
    uint64_t dims[3] = {(uint64_t)p.in_w, (uint64_t)p.in_h, (uint64_t)p.planes};
    uint64_t strides[2] = {(uint64_t)p.in_w * 4, (uint64_t)p.in_w * 4 * p.in_h};
    uint32_t box[3] = {(uint32_t)C::IN_TW, (uint32_t)C::IN_TH, 1};
    int st = encode_tensor_map(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.in, dims, strides, box, TMA_SW_NONE);
    if (st != CSD_OK) return st;
    upfirdn2d_tile_kernel<UP, DOWN, K, true><<<(unsigned)blocks, 256, 0, stream>>>(map, p);
  } else {
    upfirdn2d_tile_kernel<UP, DOWN, K, false><<<(unsigned)blocks, 256, 0, stream>>>(map, p);
  }
  CSD_LAUNCH_CHECK("upfirdn2d_tile_kernel");
  return CSD_OK;
}

// ---- NHWC bf16 FIR x2 resampling (4-tap separable filter) -----------------------------------------
// mode 1: up x2, pad (2,1), taps scaled by 2 per axis  (upsample_2d, up_or_down_sampling.py:195-224)
// mode 2: down x2, pad (1,1)                           (downsample_2d, :227-257)
// One thread = one output pixel x 8 channels (16-byte vectors, coalesced along channels).
template <typename VT>
struct FirNhwcParamsT {
  const VT* src;
  VT* out;
  const VT* add;
  int batch, h, w, oh, ow, cvec;  // cvec = channel pitch / 8
  float kf[4];                    // flipped, normalised 1-D taps (already x2 for mode 1)
  int round_out;                  // fp32 tensors: round the output to tf32 (it only feeds tensor-core operands)
  // fused GroupNorm(+SiLU) on the INPUT (TMA-staged kernel only): out = FIR(act(src * scale + shift)) with the
  // per-(image, channel) (scale, shift) table of gn_coeffs - the normalised tensor never exists in HBM
  const float2* norm;             // [batch, norm_c] or null
  int norm_c;                     // channels covered by the table (multiple of 8); channels beyond it stay as loaded
  int norm_silu;
};
using FirNhwcParams = FirNhwcParamsT<bf16x8>;

template <int MODE, typename VT>
__global__ void __launch_bounds__(256) fir_nhwc_kernel(FirNhwcParamsT<VT> p) {
  const long long total = (long long)p.batch * p.oh * p.ow * p.cvec;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % p.cvec);
    long long pix = idx / p.cvec;
    const int ox = (int)(pix % p.ow);
    const int oy = (int)((pix / p.ow) % p.oh);
    const int b = (int)(pix / ((long long)p.ow * p.oh));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // contributing input rows / columns and their weights
    int iy[4], ix[4];
    float wy[4], wx[4];
    int ny, nx;
    if (MODE == 1) {
      // out[2i] = in[i-1] kf[0] + in[i] kf[2];  out[2i+1] = in[i] kf[1] + in[i+1] kf[3]
      const int i = oy >> 1, j = ox >> 1;
      ny = nx = 2;
      if ((oy & 1) == 0) { iy[0] = i - 1; wy[0] = p.kf[0]; iy[1] = i; wy[1] = p.kf[2]; }
      else               { iy[0] = i;     wy[0] = p.kf[1]; iy[1] = i + 1; wy[1] = p.kf[3]; }
      if ((ox & 1) == 0) { ix[0] = j - 1; wx[0] = p.kf[0]; ix[1] = j; wx[1] = p.kf[2]; }
      else               { ix[0] = j;     wx[0] = p.kf[1]; ix[1] = j + 1; wx[1] = p.kf[3]; }
    } else if (MODE == 2) {
      // out[o] = sum_k kf[k] in[2o - 1 + k]
      ny = nx = 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        iy[k] = 2 * oy - 1 + k; wy[k] = p.kf[k];
        ix[k] = 2 * ox - 1 + k; wx[k] = p.kf[k];
      }
    } else {
      // same-rate filter with pad (2,2): out[o] = sum_k kf[k] in[o - 2 + k], o in [0, h]
      ny = nx = 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        iy[k] = oy - 2 + k; wy[k] = p.kf[k];
        ix[k] = ox - 2 + k; wx[k] = p.kf[k];
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (a >= ny) break;
      if (iy[a] < 0 || iy[a] >= p.h) continue;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c >= nx) break;
        if (ix[c] < 0 || ix[c] >= p.w) continue;
        const VT v = p.src[(((long long)b * p.h + iy[a]) * p.w + ix[c]) * p.cvec + cv];
        float f[8];
        unpack8(v, f);
        const float wgt = wy[a] * wx[c];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(wgt, f[i], acc[i]);
      }
    }
    if (p.add != nullptr) {
      float f[8];
      unpack8(p.add[idx], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
    if (sizeof(VT) == 32 && p.round_out) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = round_tf32(acc[i]);
    }
    p.out[idx] = pack8_as<VT>(acc);
  }
}

// ---- TMA-staged variant (modes 1 and 2) ------------------------------------------------------------------
// The per-output kernel above issues 4 (up) / 16 (down) 16-byte global loads per 16-byte store and is bound by
// the L1 tag stage (ncu on B200: l1tex throughput 92 %, DRAM 20 %, 1.0-1.7 TB/s). Here one CTA owns an output
// tile (16x16 up / 8x8 down) x one chunk of <= 64 channels: its input footprint (10x10 / 18x18 pixels) arrives in
// shared memory through ONE TMA box load over [batch, h, w, c] (pixels outside the image are zero-filled by the
// TMA unit = the FIR's zero padding), all taps are read from shared memory, and outputs leave as 16-byte vectors.
template <int MODE>
struct FirTile {
  static constexpr int TO = MODE == 1 ? 16 : 8;                 // output tile edge
  static constexpr int TI = MODE == 1 ? TO / 2 + 2 : 2 * TO + 2;  // input tile edge
};

template <int MODE, typename VT>
__global__ void __launch_bounds__(256)
fir_tma_kernel(const __grid_constant__ CUtensorMap map, const FirNhwcParamsT<VT> p, int tiles_x, int tiles_y, int cchunks,
               int cb /* channels per chunk, multiple of 8, <= 64 (bf16) / 32 (fp32) */) {
  using T = FirTile<MODE>;
  extern __shared__ __align__(128) uint8_t fir_smem[];
  __shared__ __align__(8) unsigned long long bar;
  int t = blockIdx.x;
  const int cc = t % cchunks;
  t /= cchunks;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y, b = t / tiles_y;
  const int ox0 = tx * T::TO, oy0 = ty * T::TO;
  const int ix0 = MODE == 1 ? ox0 / 2 - 1 : 2 * ox0 - 1;
  const int iy0 = MODE == 1 ? oy0 / 2 - 1 : 2 * oy0 - 1;
  const int cvs = cb >> 3;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_mbar_init();
    ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar), (uint32_t)(T::TI * T::TI * cb * (sizeof(VT) / 8)));
    ptx::tma_load_4d(ptx::smem_u32(fir_smem), &map, ptx::smem_u32(&bar), cc * cb, ix0, iy0, b);
  }
  __syncthreads();
  ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  if (p.norm != nullptr) {
    // normalise + activate the staged input tile in place, every sample once (the quad loop below reads each sample
    // up to 4 / 9 times). Pixels outside the image keep the zeros TMA wrote: the reference resamples the ACTIVATED
    // tensor with zero padding (models/layerspp.py:245-258 after act(GroupNorm_0(x))).
    VT* tl = reinterpret_cast<VT*>(fir_smem);
    const int nvec = T::TI * T::TI * cvs;
    for (int i = threadIdx.x; i < nvec; i += 256) {
      const int cv = i % cvs;
      const int px = (i / cvs) % T::TI, py = i / (cvs * T::TI);
      const int gy = iy0 + py, gx = ix0 + px;
      const int c0 = (cc * cvs + cv) * 8;
      if (gy < 0 || gy >= p.h || gx < 0 || gx >= p.w || c0 + 8 > p.norm_c) continue;
      float f[8];
      unpack8(tl[i], f);
      const float4* cf = reinterpret_cast<const float4*>(p.norm + (long long)b * p.norm_c + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t4 = __ldg(cf + j);
        float y0 = fmaf(f[2 * j], t4.x, t4.y), y1 = fmaf(f[2 * j + 1], t4.z, t4.w);
        if (p.norm_silu) {
          y0 = silu_act<VT>(y0);
          y1 = silu_act<VT>(y1);
        }
        f[2 * j] = y0;
        f[2 * j + 1] = y1;
      }
      tl[i] = pack8_as<VT>(f);
    }
    __syncthreads();
  }
  const VT* tile = reinterpret_cast<const VT*>(fir_smem);         // [TI][TI][cvs] 8-channel vectors
  // One work item = a 2 x 2 output quad x one 8-channel vector. The quad's input footprint (3 x 3 pixels for up x2,
  // 6 x 6 for down x2) is read and unpacked once and feeds all four outputs: 2.25 / 9 shared-memory vector loads and
  // unpacks per output instead of 4 / 16 (the per-output form was instruction-issue bound: SM 70-82 % busy at 25-31 % of
  // the HBM bandwidth, profiles/named_kernels_full_r2.md). Consecutive lanes take consecutive channel vectors of the
  // same quad, so every shared-memory access is a contiguous 16/32-byte-per-lane row: conflict free.
  constexpr int Q = T::TO / 2;
  const int items = Q * Q * cvs;
  for (int item = threadIdx.x; item < items; item += 256) {
    const int cv = item % cvs;
    const int qd = item / cvs;
    const int qx = qd % Q, qy = qd / Q;
    const int ox = ox0 + 2 * qx, oy = oy0 + 2 * qy;
    const int gcv = cc * cvs + cv;                                // channel vector in the tensor
    if (ox >= p.ow || oy >= p.oh || gcv >= p.cvec) continue;
    float acc[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[o][e] = 0.f;
    if (MODE == 1) {
      // out[2i] = in[i-1] kf[0] + in[i] kf[2];  out[2i+1] = in[i] kf[1] + in[i+1] kf[3]; local row of in[i-1] = i = qy
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float f[8];
          unpack8(tile[((qy + a) * T::TI + qx + c) * cvs + cv], f);
          // weights of this sample for the even / odd output row and column (0 where the tap does not reach)
          const float wy0 = a == 0 ? p.kf[0] : (a == 1 ? p.kf[2] : 0.f), wy1 = a == 0 ? 0.f : (a == 1 ? p.kf[1] : p.kf[3]);
          const float wx0 = c == 0 ? p.kf[0] : (c == 1 ? p.kf[2] : 0.f), wx1 = c == 0 ? 0.f : (c == 1 ? p.kf[1] : p.kf[3]);
          if (a < 2 && c < 2) {
            const float wgt = wy0 * wx0;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[0][e] = fmaf(wgt, f[e], acc[0][e]);
          }
          if (a < 2 && c > 0) {
            const float wgt = wy0 * wx1;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[1][e] = fmaf(wgt, f[e], acc[1][e]);
          }
          if (a > 0 && c < 2) {
            const float wgt = wy1 * wx0;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[2][e] = fmaf(wgt, f[e], acc[2][e]);
          }
          if (a > 0 && c > 0) {
            const float wgt = wy1 * wx1;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[3][e] = fmaf(wgt, f[e], acc[3][e]);
          }
        }
      }
    } else {
      // out[o] = sum_k kf[k] in[2o - 1 + k]; local index of in[2o - 1 + k] = 2 l + k; quad rows l = 2 qy, 2 qy + 1
#pragma unroll
      for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          float f[8];
          unpack8(tile[((4 * qy + a) * T::TI + 4 * qx + c) * cvs + cv], f);
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int ka = a - 2 * (o >> 1), kc = c - 2 * (o & 1);      // tap indices for output o (compile time)
            if (ka >= 0 && ka < 4 && kc >= 0 && kc < 4) {
              const float wgt = p.kf[ka] * p.kf[kc];
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[o][e] = fmaf(wgt, f[e], acc[o][e]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int oxx = ox + (o & 1), oyy = oy + (o >> 1);
      if (oxx >= p.ow || oyy >= p.oh) continue;
      const long long off = (((long long)b * p.oh + oyy) * p.ow + oxx) * p.cvec + gcv;
      if (p.add != nullptr) {
        float f[8];
        unpack8(p.add[off], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[o][e] += f[e];
      }
      if (sizeof(VT) == 32 && p.round_out) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[o][e] = round_tf32(acc[o][e]);
      }
      p.out[off] = pack8_as<VT>(acc[o]);
    }
  }
}

template <int MODE, typename VT>
static int launch_fir_tma(const FirNhwcParamsT<VT>& p, cudaStream_t stream) {
  using T = FirTile<MODE>;
  constexpr int eb = (int)sizeof(VT) / 8;        // bytes per element
  const int pitch = p.cvec * 8;
  // Channels per CTA. 128-byte chunks (64 bf16 / 32 fp32 channels) by default; where that cuts a pixel's channel row
  // into pieces that straddle 128-byte lines (96 bf16 channels = 192 bytes per pixel: every other 128-byte piece starts
  // at offset 64, and the 64 bytes left over are written later by another CTA) the up x2 kernel - output bound, four
  // stores per input sample - takes whole 96-channel rows instead, so a tile row leaves as one contiguous run.
  int cb = std::min(eb == 2 ? 64 : 32, pitch);
  static const bool whole_rows = [] { const char* e = getenv("CSD_FIR_CHUNK64"); return !(e && atoi(e) != 0); }();   // A/B
  if (MODE == 1 && whole_rows && eb == 2 && pitch % 96 == 0 && (pitch * eb) % 128 != 0) cb = 96;
  const int cchunks = ceil_div(pitch, cb);
  const int tiles_x = ceil_div(p.ow, T::TO), tiles_y = ceil_div(p.oh, T::TO);
  CUtensorMap map;
  uint64_t dims[4] = {(uint64_t)pitch, (uint64_t)p.w, (uint64_t)p.h, (uint64_t)p.batch};
  uint64_t strides[3] = {(uint64_t)pitch * eb, (uint64_t)pitch * eb * p.w, (uint64_t)pitch * eb * p.w * p.h};
  uint32_t box[4] = {(uint32_t)cb, (uint32_t)T::TI, (uint32_t)T::TI, 1};
  int st = encode_tensor_map(&map, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                             p.src, dims, strides, box, TMA_SW_NONE);
  if (st != CSD_OK) return st;
  const long long blocks = (long long)p.batch * tiles_x * tiles_y * cchunks;
  CSD_REQUIRE(blocks < (1LL << 31), "fir_resample: too many tiles");
  const size_t smem = (size_t)T::TI * T::TI * cb * eb;
  fir_tma_kernel<MODE, VT><<<(unsigned)blocks, 256, smem, stream>>>(map, p, tiles_x, tiles_y, cchunks, cb);
  CSD_LAUNCH_CHECK("fir_tma_kernel");
  return CSD_OK;
}

// ---- tap-stacked output heads: shifted sum of per-tap partial maps ------------------------------------------------
// A 3x3 convolution to a handful of channels (the 96 -> 6 output-skip heads, models/ncsnpp.py:337-352) wastes the tensor
// core either way round (N = 8 at the per-instruction floor, or 8 of 128 M rows). It is computed instead as ONE 1x1
// convolution to 9 * Cout "channels" - row t * Cout + co holds W[co, :, t] . a[pixel] for tap t at the UNSHIFTED pixel
// (54 of 128 M rows, fused GroupNorm + SiLU prologue, no halo) - followed by this pass:
//   out[b, y, x, co] = bias[co] + res[b, y, x, co] + sum_t P[b, y + t / 3 - 1, x + t % 3 - 1, t * Cout + co]
// with out-of-image taps contributing zero (the reference zero-pads the normalised activation, padding = 1).
// One thread per output pixel; every element of P is read exactly once.
__global__ void __launch_bounds__(256)
tap_shift_sum_kernel(const __nv_bfloat16* __restrict__ part, int p_pitch, int cout, const float* __restrict__ bias,
                     const __nv_bfloat16* __restrict__ res, int res_pitch, __nv_bfloat16* __restrict__ out, int out_pitch,
                     int batch, int H, int W) {
  const long long total = (long long)batch * H * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const long long b = idx / ((long long)W * H);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (i < cout && bias != nullptr) ? __ldg(bias + i) : 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const __nv_bfloat16* src = part + ((b * H + yy) * W + xx) * p_pitch + t * cout;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < cout) acc[i] += __bfloat162float(src[i]);
    }
    if (res != nullptr) {
      const __nv_bfloat16* rp = res + idx * res_pitch;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < cout) acc[i] += __bfloat162float(rp[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i >= cout) acc[i] = 0.f;
    __nv_bfloat16* op = out + idx * out_pitch;
    const bf16x8 o = pack8(acc);
    *reinterpret_cast<uint4*>(op) = *reinterpret_cast<const uint4*>(&o);
    for (int c = 8; c < out_pitch; c += 8) *reinterpret_cast<uint4*>(op + c) = make_uint4(0, 0, 0, 0);
  }
}


// The same sum from a shared-memory tile. The kernel above issues 9 * cout two-byte global loads per pixel at a 112-byte
// pixel pitch (profiles/step_launches_r2_final.md: 107 us for [64, 160, 160, 56], 0.33 of the HBM rate, LSU bound).
// Here a CTA owns 8 x 32 output pixels: the partial rows of the 10 x 34 halo tile arrive as coalesced 16-byte vectors,
// land in shared memory at an ODD word pitch (consecutive pixels hit different banks), and every thread sums its 9 taps
// from there - same order of additions as above (bias, taps 0..8, residual), so both kernels agree bit for bit.
// COUT is a template parameter: the vector count per pixel, the tap offsets and the index arithmetic of the staging loop
// are compile-time constants (with run-time divisors the staging loop alone was ~100 instructions per vector).
constexpr int kSsTH = 8, kSsTW = 32;

template <int COUT>
__global__ void __launch_bounds__(kSsTH * kSsTW)
tap_shift_sum_tile_kernel(const __nv_bfloat16* __restrict__ part, const float* __restrict__ bias,
                          const __nv_bfloat16* __restrict__ res, int res_pitch, __nv_bfloat16* __restrict__ out,
                          int out_pitch, int H, int W, int tiles_x, int tiles_y) {
  constexpr int PITCH = (9 * COUT + 7) / 8 * 8;        // channels per pixel row of `part`
  constexpr int VECS = PITCH / 8;                      // 16-byte vectors per pixel row
  constexpr int PW = (PITCH / 2) | 1;                  // shared-memory pixel pitch in words (odd)
  constexpr int HW_ = kSsTW + 2, HH_ = kSsTH + 2;
  __shared__ uint32_t ss_tile[HH_ * HW_ * PW];
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y, b = t / tiles_y;
  const int x0 = tx * kSsTW, y0 = ty * kSsTH;
  const __nv_bfloat16* img = part + (long long)b * H * W * PITCH;
#pragma unroll 2
  for (int i = threadIdx.x; i < HH_ * HW_ * VECS; i += kSsTH * kSsTW) {
    const int v = i % VECS, pp = i / VECS, px = pp % HW_, py = pp / HW_;
    const int gy = y0 - 1 + py, gx = x0 - 1 + px;
    uint4 q = make_uint4(0, 0, 0, 0);                  // taps outside the image contribute zero
    if (gy >= 0 && gy < H && gx >= 0 && gx < W)
      q = __ldg(reinterpret_cast<const uint4*>(img + ((long long)gy * W + gx) * PITCH) + v);
    uint32_t* d = ss_tile + pp * PW + v * 4;
    d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
  }
  __syncthreads();
  const int lx = threadIdx.x % kSsTW, ly = threadIdx.x / kSsTW;
  const int x = x0 + lx, y = y0 + ly;
  if (x >= W || y >= H) return;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = (i < COUT && bias != nullptr) ? __ldg(bias + i) : 0.f;
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    const int yy = y + tp / 3 - 1, xx = x + tp % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;      // (zeros anyway; keeps the additions of the kernel above)
    const uint32_t* row = ss_tile + ((ly + tp / 3) * HW_ + lx + tp % 3) * PW;
    if (COUT % 2 == 0) {                               // tap tp starts on a word boundary
      const uint32_t* src = row + (tp * COUT) / 2;
#pragma unroll
      for (int i = 0; i < COUT / 2; ++i) {
        const uint32_t wv = src[i];
        acc[2 * i] += __uint_as_float(wv << 16);
        acc[2 * i + 1] += __uint_as_float(wv & 0xffff0000u);
      }
    } else {
      const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(row) + tp * COUT;
#pragma unroll
      for (int i = 0; i < COUT; ++i) acc[i] += __bfloat162float(src[i]);
    }
  }
  const long long idx = ((long long)b * H + y) * W + x;
  if (res != nullptr) {
    const __nv_bfloat16* rp = res + idx * res_pitch;
#pragma unroll
    for (int i = 0; i < COUT; ++i) acc[i] += __bfloat162float(rp[i]);
  }
  __nv_bfloat16* op = out + idx * out_pitch;
  const bf16x8 o = pack8(acc);                         // (channels >= COUT are the zeros acc started with)
  *reinterpret_cast<uint4*>(op) = *reinterpret_cast<const uint4*>(&o);
  for (int c = 8; c < out_pitch; c += 8) *reinterpret_cast<uint4*>(op + c) = make_uint4(0, 0, 0, 0);
}

template <int COUT>
static void launch_shift_sum_tile(const void* partial, const float* bias, const void* res, int res_pitch, void* out,
                                  int out_pitch, int batch, int h, int w, cudaStream_t stream) {
  const int tiles_x = ceil_div(w, kSsTW), tiles_y = ceil_div(h, kSsTH);
  tap_shift_sum_tile_kernel<COUT><<<(unsigned)(batch * tiles_x * tiles_y), kSsTH * kSsTW, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(partial), bias, static_cast<const __nv_bfloat16*>(res), res_pitch,
      static_cast<__nv_bfloat16*>(out), out_pitch, h, w, tiles_x, tiles_y);
}

}  // namespace csd

extern "C" {

int csd_upfirdn2d_out_size(int in_h, int in_w, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                           int pad_x0, int pad_x1, int pad_y0, int pad_y1, int* out_h, int* out_w) {
  CSD_REQUIRE(out_h && out_w, "null output");
  CSD_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "up/down must be >= 1");
  const int hh = in_h * up_y + pad_y0 + pad_y1 - kh;
  const int ww = in_w * up_x + pad_x0 + pad_x1 - kw;
  CSD_REQUIRE(hh >= 0 && ww >= 0, "upfirdn2d: kernel larger than the padded input");
  *out_h = hh / down_y + 1;
  *out_w = ww / down_x + 1;
  return CSD_OK;
}

int csd_upfirdn2d_f32(const float* input, const float* kernel, float* output, int64_t planes, int in_h, int in_w,
                      int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1,
                      int pad_y0, int pad_y1, csd_stream_t stream_) {
  using namespace csd;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CSD_REQUIRE(planes >= 0 && in_h >= 1 && in_w >= 1, "upfirdn2d: bad input shape");
  CSD_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= 64, "upfirdn2d: kernel %dx%d unsupported (max 64 taps)", kh, kw);
  UpfirdnParams p;
  memset(&p, 0, sizeof(p));
  int st = csd_upfirdn2d_out_size(in_h, in_w, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1,
                                  &p.out_h, &p.out_w);
  if (st != CSD_OK) return st;
  if (planes == 0) return CSD_OK;  // empty batch: nothing to launch (pointers may be null)
  CSD_REQUIRE(input && kernel && output, "upfirdn2d: null pointer");
  p.in = input; p.kernel = kernel; p.out = output; p.planes = planes;
  p.in_h = in_h; p.in_w = in_w; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  const bool square = (up_x == up_y) && (down_x == down_y) && (kh == kw) && kh == 4;
  if (square && up_x == 2 && down_x == 1) return launch_tile<2, 1, 4>(p, stream);
  if (square && up_x == 1 && down_x == 2) return launch_tile<1, 2, 4>(p, stream);
  if (square && up_x == 1 && down_x == 1) return launch_tile<1, 1, 4>(p, stream);
  const long long total = planes * p.out_h * p.out_w;
  const int blocks = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)num_sms() * 32);
  upfirdn2d_generic_kernel<<<blocks, 256, 0, stream>>>(p);
  CSD_LAUNCH_CHECK("upfirdn2d_generic_kernel");
  return CSD_OK;
}

}  // extern "C"

namespace csd {
template <typename VT>
static int fir_resample_launch(const void* src, void* out, const void* add, int batch, int h, int w, int c_pitch, int mode,
                               const float* taps4_host, cudaStream_t stream, const float* norm = nullptr, int norm_c = 0,
                               int norm_silu = 0) {
  CSD_REQUIRE(src && out && taps4_host, "fir_resample: null pointer");
  CSD_REQUIRE(norm == nullptr || (norm_c >= 8 && norm_c % 8 == 0 && norm_c <= c_pitch),
              "fir_resample: fused GroupNorm table covers %d channels (multiple of 8, <= pitch %d)", norm_c, c_pitch);
  const int round_out = (mode & 0x10) ? 1 : 0;
  mode &= 0xf;
  CSD_REQUIRE(mode >= 1 && mode <= 3, "fir_resample: mode %d (1 = up, 2 = down, 3 = pre-filter)", mode);
  CSD_REQUIRE(c_pitch % 8 == 0, "fir_resample: channel pitch %d not a multiple of 8", c_pitch);
  CSD_REQUIRE(mode != 2 || (h % 2 == 0 && w % 2 == 0), "fir_resample: odd size %dx%d for downsampling", h, w);
  FirNhwcParamsT<VT> p;
  p.src = static_cast<const VT*>(src);
  p.out = static_cast<VT*>(out);
  p.add = static_cast<const VT*>(add);
  p.batch = batch; p.h = h; p.w = w; p.cvec = c_pitch / 8;
  p.round_out = round_out;
  p.norm = reinterpret_cast<const float2*>(norm); p.norm_c = norm_c; p.norm_silu = norm_silu;
  p.oh = mode == 1 ? h * 2 : (mode == 2 ? h / 2 : h + 1);
  p.ow = mode == 1 ? w * 2 : (mode == 2 ? w / 2 : w + 1);
  float sum = 0.f;
  for (int i = 0; i < 4; ++i) sum += taps4_host[i];
  CSD_REQUIRE(sum != 0.f, "fir_resample: taps sum to zero");
  for (int i = 0; i < 4; ++i) p.kf[i] = taps4_host[3 - i] / sum * (mode == 1 ? 2.f : 1.f);
  const long long total = (long long)batch * p.oh * p.ow * p.cvec;
  if (total == 0) return CSD_OK;
  const int blocks = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)num_sms() * 16);
  static const bool per_output = getenv("CSD_FIR_PER_OUTPUT") != nullptr;   // A/B switch
  if (mode != 3 && !per_output && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    return mode == 1 ? launch_fir_tma<1, VT>(p, stream) : launch_fir_tma<2, VT>(p, stream);
  }
  if (norm != nullptr)
    return set_error(CSD_ERR_UNSUPPORTED, "fir_resample: the fused GroupNorm input transform needs the TMA-staged kernel "
                     "(modes 1 / 2, 16-byte aligned source)");
  if (mode == 1) fir_nhwc_kernel<1, VT><<<blocks, 256, 0, stream>>>(p);
  else if (mode == 2) fir_nhwc_kernel<2, VT><<<blocks, 256, 0, stream>>>(p);
  else fir_nhwc_kernel<3, VT><<<blocks, 256, 0, stream>>>(p);
  CSD_LAUNCH_CHECK("fir_nhwc_kernel");
  return CSD_OK;
}
}  // namespace csd

extern "C" {

int csd_fir_resample_nhwc_bf16(const void* src, void* out, const void* add, int batch, int h, int w, int c_pitch,
                               int mode, const float* taps4_host, csd_stream_t stream) {
  return csd::fir_resample_launch<csd::bf16x8>(src, out, add, batch, h, w, c_pitch, mode, taps4_host,
                                               static_cast<cudaStream_t>(stream));
}
int csd_fir_resample_nhwc_f32(const void* src, void* out, const void* add, int batch, int h, int w, int c_pitch,
                              int mode, const float* taps4_host, csd_stream_t stream) {
  return csd::fir_resample_launch<csd::f32x8>(src, out, add, batch, h, w, c_pitch, mode, taps4_host,
                                              static_cast<cudaStream_t>(stream));
}

int csd_fir_norm_resample_nhwc_bf16(const void* src, void* out, const void* add, const float* norm, int norm_c,
                                    int norm_silu, int batch, int h, int w, int c_pitch, int mode,
                                    const float* taps4_host, csd_stream_t stream) {
  return csd::fir_resample_launch<csd::bf16x8>(src, out, add, batch, h, w, c_pitch, mode, taps4_host,
                                               static_cast<cudaStream_t>(stream), norm, norm_c, norm_silu);
}
int csd_fir_norm_resample_nhwc_f32(const void* src, void* out, const void* add, const float* norm, int norm_c,
                                   int norm_silu, int batch, int h, int w, int c_pitch, int mode,
                                   const float* taps4_host, csd_stream_t stream) {
  return csd::fir_resample_launch<csd::f32x8>(src, out, add, batch, h, w, c_pitch, mode, taps4_host,
                                              static_cast<cudaStream_t>(stream), norm, norm_c, norm_silu);
}

int csd_tap_shift_sum_bf16(const void* partial, int p_pitch, int cout, const float* bias, const void* res, int res_pitch,
                           void* out, int out_pitch, int batch, int h, int w, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(partial && out && batch >= 1 && h >= 1 && w >= 1, "tap_shift_sum: bad arguments");
  CSD_REQUIRE(cout >= 1 && cout <= 8 && p_pitch >= 9 * cout, "tap_shift_sum: cout=%d (1..8), partial pitch %d", cout, p_pitch);
  CSD_REQUIRE(out_pitch % 8 == 0 && out_pitch >= 8 && (res == nullptr || res_pitch >= cout),
              "tap_shift_sum: output pitch %d must be a multiple of 8", out_pitch);
  static const bool legacy = [] { const char* e = getenv("CSD_SHIFT_SUM_LEGACY"); return e && atoi(e) != 0; }();   // A/B
  const long long tiles = (long long)batch * ceil_div(h, kSsTH) * ceil_div(w, kSsTW);
  // (cout = 8 would need 50 KB for the tile: it stays on the kernel above)
  if (!legacy && cout <= 7 && p_pitch == (9 * cout + 7) / 8 * 8 && tiles < (1LL << 31) && reinterpret_cast<uintptr_t>(partial) % 16 == 0) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (cout) {
      case 1: launch_shift_sum_tile<1>(partial, bias, res, res_pitch, out, out_pitch, batch, h, w, st); break;
      case 2: launch_shift_sum_tile<2>(partial, bias, res, res_pitch, out, out_pitch, batch, h, w, st); break;
      case 3: launch_shift_sum_tile<3>(partial, bias, res, res_pitch, out, out_pitch, batch, h, w, st); break;
      case 4: launch_shift_sum_tile<4>(partial, bias, res, res_pitch, out, out_pitch, batch, h, w, st); break;
      case 5: launch_shift_sum_tile<5>(partial, bias, res, res_pitch, out, out_pitch, batch, h, w, st); break;
      case 6: launch_shift_sum_tile<6>(partial, bias, res, res_pitch, out, out_pitch, batch, h, w, st); break;
      case 7: launch_shift_sum_tile<7>(partial, bias, res, res_pitch, out, out_pitch, batch, h, w, st); break;
      default: break;
    }
    CSD_LAUNCH_CHECK("tap_shift_sum_tile_kernel");
    return CSD_OK;
  }
  const long long total = (long long)batch * h * w;
  const int blocks = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)num_sms() * 16);
  tap_shift_sum_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(partial), p_pitch, cout, bias, static_cast<const __nv_bfloat16*>(res), res_pitch,
      static_cast<__nv_bfloat16*>(out), out_pitch, batch, h, w);
  CSD_LAUNCH_CHECK("tap_shift_sum_kernel");
  return CSD_OK;
}

}  // extern "C"
