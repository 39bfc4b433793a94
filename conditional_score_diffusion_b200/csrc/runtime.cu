// Library-wide runtime bits: error text, device queries, tensor-map encoder lookup.
#include "common.cuh"
#include "tensormap.cuh"
#include "../../include/csd_b200.h"

#include <atomic>
#include <cstring>
#include <mutex>

namespace csd {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

char* error_buffer() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int encode_tensor_map(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      TmaSwizzle swizzle, const uint32_t* elem_strides) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return set_error(CSD_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  CSD_REQUIRE(rank >= 1 && rank <= 5, "tensor map rank %d out of range", rank);
  CSD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base %p not 16-byte aligned", base);
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    CSD_REQUIRE(box[i] >= 1 && box[i] <= 256, "tensor map box[%d]=%u out of range", i, box[i]);
    CSD_REQUIRE(dims[i] >= 1, "tensor map dim[%d]=0", i);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gs[i] = strides_bytes[i];
    CSD_REQUIRE((strides_bytes[i] & 15) == 0, "tensor map stride[%d]=%llu not a multiple of 16", i,
                (unsigned long long)strides_bytes[i]);
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle == TMA_SW_32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  if (swizzle == TMA_SW_64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  if (swizzle == TMA_SW_128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(CSD_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu,%llu] "
                     "box=[%u,%u,%u,%u,%u]",
                     (int)r, rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
                     (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
                     (unsigned long long)(rank > 4 ? gd[4] : 0), bx[0], rank > 1 ? bx[1] : 0,
                     rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
  }
  return CSD_OK;
}

}  // namespace csd

extern "C" {

const char* csd_last_error(void) { return csd::error_buffer(); }

int csd_abi_version(void) { return CSD_ABI_VERSION; }

long long csd_launch_count(void) { return csd::g_launches.load(std::memory_order_relaxed); }

int csd_device_sm_count(int* out) {
  if (!out) return csd::set_error(CSD_ERR_INVALID, "null output");
  int dev = 0;
  CSD_CUDA(cudaGetDevice(&dev));
  *out = csd::num_sms();
  return CSD_OK;
}

}  // extern "C"
