// Batched weight (re-)packing: fp32 parameters in the reference's layouts -> the bf16 K-major operand images of
// csd_conv_gemm, for the forward convolutions, the data-gradient convolutions (flipped taps, transposed channels)
// and the bias vectors, all in ONE launch driven by a device-resident job table.
//
// The reference keeps fp32 nn.Parameters and lets cuDNN pick layouts per call; here every optimizer step is followed
// by one csd_pack_weights call that refreshes all ~400 packed operands of the network in place (the recorded launch
// lists and their CUDA graphs keep pointing at the same buffers).
#include "common.cuh"
#include "../../include/csd_b200.h"

namespace csd {

__global__ void __launch_bounds__(256) pack_weights_kernel(const csd_pack_job* __restrict__ jobs) {
  const csd_pack_job j = jobs[blockIdx.y];
  const long long total = (long long)j.rows * j.cols * j.taps;
  if (j.kind == 1) {   // fp32 vector: dst[i] = src[i] + src2[i]
    float* d = static_cast<float*>(j.dst);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < j.rows; i += (long long)gridDim.x * blockDim.x)
      d[i] = j.src[i] + (j.src2 != nullptr ? j.src2[i] : 0.f);
    return;
  }
  __nv_bfloat16* d = static_cast<__nv_bfloat16*>(j.dst);
  float* df = static_cast<float*>(j.dst);     // kind 2: the same image in fp32 (tf32 plan)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // destination order (row, tap, col): consecutive threads write consecutive bf16 elements
    const int c = (int)(i % j.cols);
    const int tap = (int)((i / j.cols) % j.taps);
    const int r = (int)(i / ((long long)j.cols * j.taps));
    const int ts = j.flip ? j.taps - 1 - tap : tap;
    const float v = j.src[(long long)r * j.s_row + (long long)c * j.s_col + (long long)ts * j.s_tap] * j.scale;
    const long long o = (long long)r * j.dst_pitch + (long long)tap * j.k_pad + c;
    if (j.kind == 2) df[o] = round_tf32(v);
    else d[o] = __float2bfloat16_rn(v);
  }
}

}  // namespace csd

extern "C" int csd_pack_weights(const csd_pack_job* jobs_dev, int njobs, int64_t max_elems, csd_stream_t stream) {
  using namespace csd;
  CSD_REQUIRE(jobs_dev != nullptr && njobs >= 1 && njobs <= 65535 && max_elems >= 1, "pack_weights: bad arguments");
  const int bx = (int)std::max<long long>(1, std::min<long long>(ceil_div_ll(max_elems, 256 * 4), 64));
  pack_weights_kernel<<<dim3((unsigned)bx, (unsigned)njobs), 256, 0, static_cast<cudaStream_t>(stream)>>>(jobs_dev);
  CSD_LAUNCH_CHECK("pack_weights_kernel");
  return CSD_OK;
}
