// fq_rows.cuh - one warp processes one row segment whose qparams are uniform (per-channel
// quantisation).  Shared by the single-tensor row kernels and the grouped (multi-tensor) launch.
#pragma once
#include "fq_math.cuh"

namespace dlmcq {

// Forward over `len` contiguous elements starting at xr.  128-bit accesses when the segment start
// is 16-byte aligned in every tensor, two loads in flight per lane; scalar otherwise / for the tail.
template <int FORM, typename T>
__device__ __forceinline__ void fwd_row_segment(const T* __restrict__ xr, T* __restrict__ yr, T* __restrict__ cr,
                                                int64_t len, const ChanParams& p, float lo, float hi, int lane) {
  using V = Vec<T>;
  using raw = typename V::raw;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(yr) |
                        reinterpret_cast<uintptr_t>(cr)) & 15u) == 0;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t nvec = len / V::N;
    const raw* xv = reinterpret_cast<const raw*>(xr);
    constexpr int U = 4;                       // 128-bit loads in flight per lane
    for (int64_t j = lane; j < nvec; j += 32 * U) {
      raw r[U];
#pragma unroll
      for (int h = 0; h < U; ++h)
        if (j + 32 * h < nvec) r[h] = ld_stream(xv + j + 32 * h);
#pragma unroll
      for (int h = 0; h < U; ++h) {
        if (j + 32 * h < nvec) {
          float f[V::N], fy[V::N], fc[V::N];
          V::unpack(r[h], f);
          fq_vec<FORM, V::N>(f, p, lo, hi, fc, fy);
          if (yr) st_stream(reinterpret_cast<raw*>(yr) + j + 32 * h, V::pack(fy));
          if (cr) st_stream(reinterpret_cast<raw*>(cr) + j + 32 * h, V::pack(fc));
        }
      }
    }
    done = nvec * V::N;
  }
  for (int64_t j = done + lane; j < len; j += 32) {
    float c, v;
    fq_elem<FORM>(to_f32<T>(xr[j]), p, lo, hi, c, v);
    if (yr) yr[j] = from_f32<T>(v);
    if (cr) cr[j] = from_f32<T>(c);
  }
}

// Backward over one row segment: writes dx, returns the warp-reduced scale / offset terms
// (valid in every lane).
template <int FORM, typename T>
__device__ __forceinline__ void bwd_row_segment(const T* __restrict__ xr, const T* __restrict__ gr,
                                                T* __restrict__ dr, int64_t len, const ChanParams& p, float lo,
                                                float hi, int lane, float& as, float& ao) {
  using V = Vec<T>;
  using raw = typename V::raw;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(gr) |
                        reinterpret_cast<uintptr_t>(dr)) & 15u) == 0;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t nvec = len / V::N;
    const raw* xv = reinterpret_cast<const raw*>(xr);
    const raw* gv = reinterpret_cast<const raw*>(gr);
    for (int64_t j = lane; j < nvec; j += 64) {
      const bool two = (j + 32) < nvec;
      raw x0 = ld_stream(xv + j), g0 = ld_stream(gv + j), x1 = x0, g1 = g0;
      if (two) { x1 = ld_stream(xv + j + 32); g1 = ld_stream(gv + j + 32); }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        float fx[V::N], fg[V::N], fo[V::N];
        V::unpack(h ? x1 : x0, fx);
        V::unpack(h ? g1 : g0, fg);
        fq_vec_bwd<FORM, true, V::N>(fx, fg, p, lo, hi, fo, as, ao);
        st_stream(reinterpret_cast<raw*>(dr) + j + 32 * h, V::pack(fo));
      }
    }
    done = nvec * V::N;
  }
  for (int64_t j = done + lane; j < len; j += 32)
    dr[j] = from_f32<T>(fq_elem_bwd<FORM, true>(to_f32<T>(xr[j]), to_f32<T>(gr[j]), p, lo, hi, as, ao));
  as = warp_sum(as);
  ao = warp_sum(ao);
}

}  // namespace dlmcq
