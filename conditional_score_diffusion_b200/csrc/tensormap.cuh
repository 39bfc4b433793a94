// Host-side TMA tensor-map encoding. libcuda is never linked: the encoder entry point is
// resolved at run time through the CUDA runtime (cudaGetDriverEntryPoint), so the library
// still loads on a machine without a driver (the CPU-only symbol test relies on that).
#pragma once

#include <cuda.h>
#include "common.cuh"

namespace csd {

enum TmaSwizzle { TMA_SW_NONE = 0, TMA_SW_32 = 1, TMA_SW_64 = 2, TMA_SW_128 = 3 };

// rank <= 5. dims/box in elements (dim 0 fastest), strides in BYTES for dims 1..rank-1.
int encode_tensor_map(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      TmaSwizzle swizzle, const uint32_t* elem_strides = nullptr);

}  // namespace csd
