// misc_kernels.cu - the small helpers of dlmc/quantization/scalar/utils.py and RootQ/function.py that
// the reference exports as free functions (round_pass, floor_pass, sgn, grad_scale).  The fused
// kernels do not call these; they exist so that the functional API surface is complete on device.
#include "fq_math.cuh"

namespace dlmcq {

// mode 0: round_pass value (utils.py:29-32)   (round(x) - x) + x
// mode 1: floor_pass value (utils.py:34-37)   (floor(x) - x) + x
// mode 2: sgn              (RootQ/function.py:5-8) sign(x), sign(NaN) = 0
template <typename T>
__global__ void __launch_bounds__(kThreads)
ste_value_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, int mode) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = to_f32<T>(x[i]);
    float r;
    if (mode == 0) r = round_pass(v);
    else if (mode == 1) { const float f = floorf(v); r = (f - v) + v; }
    else r = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
    y[i] = from_f32<T>(r);
  }
}

// utils.py:24-27 grad_scale value: (s - s*g) + s*g
__global__ void grad_scale_kernel(const float* __restrict__ s, float* __restrict__ out, int64_t n, float g) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const float sg = s[i] * g;
    out[i] = (s[i] - sg) + sg;
  }
}

}  // namespace dlmcq

using namespace dlmcq;

extern "C" int dlmcq_ste_value(const void* x, void* y, int64_t numel, int dtype, int mode, void* stream) {
  if (!x || !y || numel < 0 || mode < 0 || mode > 2) return DLMCQ_EINVAL;
  if (numel == 0) return DLMCQ_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = stream_grid((numel + kThreads - 1) / kThreads, 8);
  if (dtype == DLMCQ_F32)
    ste_value_kernel<float><<<grid, kThreads, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(y), numel, mode);
  else if (dtype == DLMCQ_BF16)
    ste_value_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                               static_cast<__nv_bfloat16*>(y), numel, mode);
  else
    return DLMCQ_EINVAL;
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

extern "C" int dlmcq_grad_scale_value(const float* s, float* out, int64_t numel, float g, void* stream) {
  if (!s || !out || numel < 0) return DLMCQ_EINVAL;
  if (numel == 0) return DLMCQ_OK;
  grad_scale_kernel<<<static_cast<unsigned>((numel + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(s, out, numel, g);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}
