// Shared host/device helpers for libcsd_b200 (sm_100a only).
//
// Everything in csrc/ is reached through the C ABI declared in include/csd_b200.h:
// plain pointers and sizes in, int status out, caller-allocated outputs, caller's stream.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#define CSD_OK 0
#define CSD_ERR_INVALID 1   // bad argument (shape, alignment, null pointer)
#define CSD_ERR_CUDA 2      // CUDA runtime/driver error (see csd_last_error)
#define CSD_ERR_UNSUPPORTED 3

namespace csd {

// Thread-local last-error text, returned by csd_last_error().
char* error_buffer();
int set_error(int code, const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return CSD_OK;
  return set_error(CSD_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CSD_CUDA(expr)                                         \
  do {                                                         \
    int _s = ::csd::check_cuda((expr), #expr);                 \
    if (_s != CSD_OK) return _s;                               \
  } while (0)

#define CSD_REQUIRE(cond, ...)                                 \
  do {                                                         \
    if (!(cond)) return ::csd::set_error(CSD_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Every kernel launch in the library goes through this macro: it checks the launch and counts it
// (csd_launch_count), so callers can report how many of OUR kernels a region launched.
#define CSD_LAUNCH_CHECK(name)            \
  do {                                    \
    ::csd::count_launch();                \
    CSD_CUDA(cudaGetLastError());         \
  } while (0)

#ifdef __CUDACC__
#define CSD_HD __host__ __device__
#else
#define CSD_HD
#endif
CSD_HD inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
CSD_HD inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

int num_sms();  // SM count of the current device (cached)
void count_launch();

#ifdef __CUDACC__

// SiLU with the fast-division path (2 MUFU + 2 FP ops); |error| < 2 ulp of fp32, far below bf16 storage.
__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// SiLU through one MUFU op: x*sigmoid(x) = h + h*tanh(h) with h = x/2 (tanh.approx.f32, max rel. error
// 2^-11: below the bf16 rounding applied to the result). Used where MUFU throughput is the constraint.
__device__ __forceinline__ float silu_tanh(float v) {
  const float h = 0.5f * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 16-byte vector of 8 bf16 values.
struct __align__(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}

// 32-byte vector of 8 fp32 values: the fp32-activation ("tf32" precision) counterpart of bf16x8. Kernels that walk
// NHWC activations are templates over the vector type; channel pitches are multiples of 8, so both are aligned.
struct __align__(32) f32x8 {
  float4 lo, hi;
};

__device__ __forceinline__ void unpack8(const f32x8& p, float* f) {
  f[0] = p.lo.x; f[1] = p.lo.y; f[2] = p.lo.z; f[3] = p.lo.w;
  f[4] = p.hi.x; f[5] = p.hi.y; f[6] = p.hi.z; f[7] = p.hi.w;
}

template <typename VT> __device__ __forceinline__ VT pack8_as(const float* f);
template <> __device__ __forceinline__ bf16x8 pack8_as<bf16x8>(const float* f) { return pack8(f); }
template <> __device__ __forceinline__ f32x8 pack8_as<f32x8>(const float* f) {
  f32x8 p;
  p.lo = make_float4(f[0], f[1], f[2], f[3]);
  p.hi = make_float4(f[4], f[5], f[6], f[7]);
  return p;
}

// SiLU per storage precision: the fast-division form for bf16 storage (error far below the bf16 rounding of the
// result), the full-precision form for fp32 storage (the "tf32" plan is held to 1e-3 against the fp32 oracle).
template <typename VT> __device__ __forceinline__ float silu_act(float v);
template <> __device__ __forceinline__ float silu_act<bf16x8>(float v) { return silu_f(v); }
template <> __device__ __forceinline__ float silu_act<f32x8>(float v) { return v / (1.0f + expf(-v)); }

// Round-to-nearest fp32 -> tf32 (10 explicit mantissa bits, low 13 bits zero). tcgen05 kind::tf32 TRUNCATES the fp32
// words it reads, which biases every product towards zero by ~2^-11; tensors of the tf32 plan that are consumed only as
// tensor-core operands (normalised activations, FIR outputs, attention q/k/v/probabilities, packed weights) are
// therefore rounded when they are produced - what cuDNN/CUTLASS TF32 convolutions do to their operands internally.
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// scalar element type of a vector type
template <typename VT> struct ElemOf;
template <> struct ElemOf<bf16x8> { using type = __nv_bfloat16; };
template <> struct ElemOf<f32x8> { using type = float; };
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }

// Streaming 16-byte global load/store (read-once data: keep it out of L1).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

#endif  // __CUDACC__

}  // namespace csd
