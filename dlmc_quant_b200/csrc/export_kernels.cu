// export_kernels.cu - true integer export of the quantised codes.
//
// The reference never materialises integers: "codes" exist only as fp32-valued numbers inside the eager
// chain, and post_training_quantization.py:95-101 saves fp32 state.  These kernels emit the codes of any
// of the four forms as int8 / uint8, or packed two-per-byte for <= 4-bit ranges (low nibble first, two's
// complement for signed ranges), and read them back (unpack + dequantise) bit-identically to the
// fake-quant forward: 4 B read + 1 or 0.5 B written per element.
#include "fq_math.cuh"

namespace dlmcq {

template <int FORM, typename T, bool PACK4>
__global__ void __launch_bounds__(kThreads)
export_codes_kernel(const T* __restrict__ x, uint8_t* __restrict__ out, int64_t n, int64_t channels, int64_t inner,
                    const float* __restrict__ scale, const float* __restrict__ offset, float g, float lo, float hi) {
  // each thread owns 8 consecutive elements -> 8 bytes (int8) or 4 bytes (packed int4)
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t groups = (n + 7) / 8;
  for (int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; gi < groups; gi += stride) {
    const int64_t i0 = gi * 8;
    uint32_t bytes[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int64_t i = i0 + e;
      float code = 0.f, y;
      if (i < n) {
        const int64_t ch = channels == 1 ? 0 : (i / inner) % channels;
        const ChanParams p = make_params<FORM>(scale, offset, ch, g, lo, hi);
        fq_elem_ref<FORM>(to_f32<T>(x[i]), p, lo, hi, code, y);       // literal chain: export is a one-off
      }
      const int c = (code == code) ? static_cast<int>(code) : 0;      // NaN code -> 0
      bytes[e] = static_cast<uint32_t>(c) & (PACK4 ? 0xFu : 0xFFu);
    }
    if (PACK4) {
      uint32_t w = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) w |= bytes[e] << (4 * e);
      const int64_t ob = gi * 4;
      const int64_t nb = (n + 1) / 2;
      for (int b = 0; b < 4; ++b)
        if (ob + b < nb) out[ob + b] = static_cast<uint8_t>(w >> (8 * b));
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (i0 + e < n) out[i0 + e] = static_cast<uint8_t>(bytes[e]);
    }
  }
}

template <int FORM, typename T, bool PACK4>
__global__ void __launch_bounds__(kThreads)
import_codes_kernel(const uint8_t* __restrict__ in, T* __restrict__ y, int64_t n, int64_t channels, int64_t inner,
                    const float* __restrict__ scale, const float* __restrict__ offset, float g, float lo, float hi,
                    int is_signed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    int c;
    if (PACK4) {
      const uint32_t nib = (in[i >> 1] >> (4 * (i & 1))) & 0xFu;
      c = is_signed ? (static_cast<int>(nib << 28) >> 28) : static_cast<int>(nib);
    } else {
      c = is_signed ? static_cast<int>(static_cast<int8_t>(in[i])) : static_cast<int>(in[i]);
    }
    const int64_t ch = channels == 1 ? 0 : (i / inner) % channels;
    const ChanParams p = make_params<FORM>(scale, offset, ch, g, lo, hi);
    const float code = static_cast<float>(c);
    float v;
    if (FORM == DLMCQ_FORM_A1 || FORM == DLMCQ_FORM_AFFINE) v = code * p.mul + p.off;
    else if (FORM == DLMCQ_FORM_ZP) v = (code - p.off) * p.mul;
    else v = code * p.mul;
    y[i] = from_f32<T>(v);
  }
}

template <typename T, bool PACK4>
static int export_dispatch(const void* x, void* out, const dlmcq_layout* l, const dlmcq_qparams* qp, cudaStream_t st) {
  const int64_t n = l->outer * l->channels * l->inner;
  const int grid = stream_grid(((n + 7) / 8 + kThreads - 1) / kThreads, 8);
  const float lo = static_cast<float>(qp->lo), hi = static_cast<float>(qp->hi);
  const T* xp = static_cast<const T*>(x);
  uint8_t* op = static_cast<uint8_t*>(out);
#define DLMCQ_EXP(F) export_codes_kernel<F, T, PACK4><<<grid, kThreads, 0, st>>>(xp, op, n, l->channels, l->inner, \
                                                                                qp->scale, qp->offset, qp->g, lo, hi)
  switch (qp->form) {
    case DLMCQ_FORM_A1: DLMCQ_EXP(DLMCQ_FORM_A1); break;
    case DLMCQ_FORM_AFFINE: DLMCQ_EXP(DLMCQ_FORM_AFFINE); break;
    case DLMCQ_FORM_ZP: DLMCQ_EXP(DLMCQ_FORM_ZP); break;
    case DLMCQ_FORM_SYM: DLMCQ_EXP(DLMCQ_FORM_SYM); break;
    default: return DLMCQ_EINVAL;
  }
#undef DLMCQ_EXP
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

template <typename T, bool PACK4>
static int import_dispatch(const void* in, void* y, const dlmcq_layout* l, const dlmcq_qparams* qp, cudaStream_t st) {
  const int64_t n = l->outer * l->channels * l->inner;
  const int grid = stream_grid((n + kThreads - 1) / kThreads, 8);
  const float lo = static_cast<float>(qp->lo), hi = static_cast<float>(qp->hi);
  const uint8_t* ip = static_cast<const uint8_t*>(in);
  T* yp = static_cast<T*>(y);
  const int sg = qp->lo < 0 ? 1 : 0;
#define DLMCQ_IMP(F) import_codes_kernel<F, T, PACK4><<<grid, kThreads, 0, st>>>(ip, yp, n, l->channels, l->inner, \
                                                                                qp->scale, qp->offset, qp->g, lo, hi, sg)
  switch (qp->form) {
    case DLMCQ_FORM_A1: DLMCQ_IMP(DLMCQ_FORM_A1); break;
    case DLMCQ_FORM_AFFINE: DLMCQ_IMP(DLMCQ_FORM_AFFINE); break;
    case DLMCQ_FORM_ZP: DLMCQ_IMP(DLMCQ_FORM_ZP); break;
    case DLMCQ_FORM_SYM: DLMCQ_IMP(DLMCQ_FORM_SYM); break;
    default: return DLMCQ_EINVAL;
  }
#undef DLMCQ_IMP
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

static int check_export(const dlmcq_layout* l, const dlmcq_qparams* qp, int pack4) {
  if (!l || !qp || !qp->scale || l->outer < 1 || l->channels < 1 || l->inner < 0) return DLMCQ_EINVAL;
  if (l->dtype != DLMCQ_F32 && l->dtype != DLMCQ_BF16) return DLMCQ_EINVAL;
  const int span = qp->hi - qp->lo;
  if (qp->lo < -128 || qp->hi > 255 || span > 255 || (qp->lo < 0 && qp->hi > 127)) return DLMCQ_EUNSUPPORTED;
  if (pack4 && (span > 15 || qp->lo < -8 || qp->hi > 15 || (qp->lo < 0 && qp->hi > 7))) return DLMCQ_EUNSUPPORTED;
  return DLMCQ_OK;
}

}  // namespace dlmcq

using namespace dlmcq;

extern "C" int dlmcq_export_codes(const void* x, void* out, const dlmcq_layout* layout, const dlmcq_qparams* qp,
                                  int pack4, void* stream) {
  if (int e = check_export(layout, qp, pack4)) return e;
  if (layout->outer * layout->channels * layout->inner == 0) return DLMCQ_OK;
  if (!x || !out) return DLMCQ_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout->dtype == DLMCQ_F32)
    return pack4 ? export_dispatch<float, true>(x, out, layout, qp, st) : export_dispatch<float, false>(x, out, layout, qp, st);
  return pack4 ? export_dispatch<__nv_bfloat16, true>(x, out, layout, qp, st)
               : export_dispatch<__nv_bfloat16, false>(x, out, layout, qp, st);
}

extern "C" int dlmcq_import_codes(const void* codes, void* y, const dlmcq_layout* layout, const dlmcq_qparams* qp,
                                  int pack4, void* stream) {
  if (int e = check_export(layout, qp, pack4)) return e;
  if (layout->outer * layout->channels * layout->inner == 0) return DLMCQ_OK;
  if (!codes || !y) return DLMCQ_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout->dtype == DLMCQ_F32)
    return pack4 ? import_dispatch<float, true>(codes, y, layout, qp, st) : import_dispatch<float, false>(codes, y, layout, qp, st);
  return pack4 ? import_dispatch<__nv_bfloat16, true>(codes, y, layout, qp, st)
               : import_dispatch<__nv_bfloat16, false>(codes, y, layout, qp, st);
}
