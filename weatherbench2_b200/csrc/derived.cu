// Derived variables computed on the device before the metric kernels (sm_100a).
//
// wb2_wind_speed replaces WindSpeed.compute (weatherbench2/derived_variables.py:
// 77-99): sqrt(u**2 + v**2), element by element, with the same three
// correctly rounded float32 operations NumPy performs (no FMA contraction), so
// the result is bit-identical.  8 B read + 4 B written per cell; the point is
// that the derived field stays in HBM for K1 instead of being built on the host.
#include <algorithm>

#include "common.cuh"

namespace wb2 {

__global__ void __launch_bounds__(256) wind_speed_kernel(const float* __restrict__ u,
                                                         const float* __restrict__ v,
                                                         float* __restrict__ out, int64_t n) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a = ldg_stream(u + i), b = ldg_stream(v + i);
    out[i] = __fsqrt_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
  }
}

}  // namespace wb2

using namespace wb2;

extern "C" int wb2_wind_speed(wb2_ctx* ctx, const float* u, const float* v, float* out,
                              int64_t n) {
  WB2_NVTX("wb2_wind_speed");
  WB2_REQUIRE(ctx != nullptr, "ctx is NULL");
  WB2_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return WB2_OK;
  WB2_REQUIRE(u && v && out, "NULL argument");
  DeviceGuard guard(ctx->device);
  const int64_t blocks = std::min<int64_t>((n + 255) / 256, int64_t(ctx->num_sms) * 16);
  wind_speed_kernel<<<static_cast<unsigned>(blocks), 256, 0, ctx->stream>>>(u, v, out, n);
  WB2_CUDA_TRY(cudaGetLastError());
  ctx->launches += 1;
  return WB2_OK;
}
