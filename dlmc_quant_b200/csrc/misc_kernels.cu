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

// Self-test of the residual-corrected division used by the fast path (fq_math.cuh): draws
// pseudo-random (x, s) pairs inside the fast-path domain - random sign/exponent/mantissa plus
// mantissas of the adversarial kinds (all ones, one bit, near powers of two) - and counts the
// pairs for which fast_div differs bitwise from the IEEE quotient __fdiv_rn.
__device__ __forceinline__ uint32_t mix32(uint64_t& st) {
  st = st * 6364136223846793005ull + 1442695040888963407ull;
  uint32_t x = static_cast<uint32_t>(st >> 32);
  x ^= x >> 15; x *= 0x2c1b3c6du; x ^= x >> 12; x *= 0x297a2d39u; x ^= x >> 15;
  return x;
}
__device__ __forceinline__ float craft(uint32_t r, uint32_t kind, int emin, int emax) {
  uint32_t mant = r & 0x7fffffu;
  switch (kind & 7u) {
    case 1: mant = 0x7fffffu; break;                 // all ones
    case 2: mant = 0; break;                         // power of two
    case 3: mant = 1u << (r % 23u); break;           // single bit
    case 4: mant = 0x7fffffu ^ (1u << (r % 23u)); break;
    case 5: mant &= 0x7ff000u; break;                // short mantissa (bf16-like data)
    default: break;
  }
  const int e = emin + static_cast<int>((r >> 23) % static_cast<uint32_t>(emax - emin + 1));
  const uint32_t bits = ((r >> 31) << 31) | (static_cast<uint32_t>(e + 127) << 23) | mant;
  return __uint_as_float(bits);
}
__global__ void __launch_bounds__(kThreads)
selftest_fastdiv_kernel(uint64_t seed, int per_thread, int narrow, unsigned long long* mismatches) {
  uint64_t st = seed + 0x9e3779b97f4a7c15ull * (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x + 1);
  unsigned long long bad = 0;
  for (int k = 0; k < per_thread; ++k) {
    const uint32_t a = mix32(st), b = mix32(st), c = mix32(st);
    // narrow: exponents typical of activations / scales; wide: the whole fast-path domain
    const float s = craft(a, c, narrow ? -12 : -40, narrow ? 4 : 39);
    const FastDiv d = make_fastdiv(s);
    float x = craft(b, c >> 3, narrow ? -20 : -100, narrow ? 8 : 59);
    if ((c >> 28) == 0) x = 0.f;
    const float q = fast_div(x, d), ref = __fdiv_rn(x, s);
    // Quotients below 2^-100 (possible only with |x| < 2^-60) may lose their last bits to underflow in
    // the residual; every form rounds them to code 0, so there the check is "tiny stays tiny".
    if (d.ok) {
      if (fabsf(ref) >= 0x1p-100f) { if (__float_as_uint(q) != __float_as_uint(ref)) ++bad; }
      else if (!(fabsf(q) <= 0x1p-99f)) ++bad;
    }
    // integer-times-scale numerators sit exactly on rounding ties of the quotient
    const float xt = __fmul_rn(static_cast<float>(static_cast<int>(b % 511u) - 255) + 0.5f, s);
    const float qt = fast_div(xt, d), rt = __fdiv_rn(xt, s);
    if (d.ok && __float_as_uint(qt) != __float_as_uint(rt) && !(qt == 0.f && rt == 0.f)) ++bad;
    // quotients that are exact integers or exact ties must come out exact (they decide the codes)
    const float xi = __fmul_rn(static_cast<float>(static_cast<int>(a % 65535u) - 32767), s);
    const float qi = fast_div(xi, d), ri = __fdiv_rn(xi, s);
    if (d.ok && __float_as_uint(qi) != __float_as_uint(ri) && !(qi == 0.f && ri == 0.f)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

}  // namespace dlmcq

using namespace dlmcq;

extern "C" int dlmcq_selftest_fastdiv(uint64_t seed, int blocks, int per_thread, int narrow,
                                      unsigned long long* mismatches_dev, void* stream) {
  if (!mismatches_dev || blocks < 1 || per_thread < 1) return DLMCQ_EINVAL;
  selftest_fastdiv_kernel<<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(seed, per_thread, narrow,
                                                                                     mismatches_dev);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

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
