// fq_rows.cuh - one warp processes one row segment whose qparams are uniform (per-channel
// quantisation).  Shared by the single-tensor row kernels and the grouped (multi-tensor) launch.
#pragma once
#include "fq_math.cuh"

namespace dlmcq {

// Forward over `len` contiguous elements starting at xr.  128-bit accesses when the segment start
// is 16-byte aligned in every tensor, U loads in flight per lane; scalar otherwise / for the tail.
template <int FORM, typename T, int U = 4>
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
    // U 128-bit loads in flight per lane: 4 in the grouped (multi-tensor) launch, 8 in the single-tensor
    // kernel, whose long activation rows profit (measured 5.76 -> 5.97 TB/s on [334,64,56,56])
    auto one = [&](const raw& r, int64_t idx) {
      float f[V::N], fy[V::N], fc[V::N];
      V::unpack(r, f);
      fq_vec<FORM, V::N>(f, p, lo, hi, fc, fy);
      if (yr) st_stream(reinterpret_cast<raw*>(yr) + idx, V::pack(fy));
      if (cr) st_stream(reinterpret_cast<raw*>(cr) + idx, V::pack(fc));
    };
    int64_t j = lane;
    for (; j + 32 * (U - 1) < nvec; j += 32 * U) {
      raw r[U];
#pragma unroll
      for (int h = 0; h < U; ++h) r[h] = ld_stream(xv + j + 32 * h);
#pragma unroll
      for (int h = 0; h < U; ++h) one(r[h], j + 32 * h);
    }
    if (j < nvec) {                            // last, partial block: still all loads first
      raw r[U];
#pragma unroll
      for (int h = 0; h < U; ++h)
        if (j + 32 * h < nvec) r[h] = ld_stream(xv + j + 32 * h);
#pragma unroll
      for (int h = 0; h < U; ++h)
        if (j + 32 * h < nvec) one(r[h], j + 32 * h);
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
    constexpr int U = 4;                       // 2 x 4 128-bit loads in flight per lane
    auto one = [&](const raw& rx, const raw& rg, int64_t idx) {
      float fx[V::N], fg[V::N], fo[V::N];
      V::unpack(rx, fx);
      V::unpack(rg, fg);
      fq_vec_bwd<FORM, true, V::N>(fx, fg, p, lo, hi, fo, as, ao);
      st_stream(reinterpret_cast<raw*>(dr) + idx, V::pack(fo));
    };
    int64_t j = lane;
    for (; j + 32 * (U - 1) < nvec; j += 32 * U) {
      raw rx[U], rg[U];
#pragma unroll
      for (int h = 0; h < U; ++h) {
        rx[h] = ld_stream(xv + j + 32 * h);
        rg[h] = ld_stream(gv + j + 32 * h);
      }
#pragma unroll
      for (int h = 0; h < U; ++h) one(rx[h], rg[h], j + 32 * h);
    }
    if (j < nvec) {                            // last, partial block: still all loads first
      raw rx[U], rg[U];
#pragma unroll
      for (int h = 0; h < U; ++h) {
        if (j + 32 * h < nvec) {
          rx[h] = ld_stream(xv + j + 32 * h);
          rg[h] = ld_stream(gv + j + 32 * h);
        }
      }
#pragma unroll
      for (int h = 0; h < U; ++h)
        if (j + 32 * h < nvec) one(rx[h], rg[h], j + 32 * h);
    }
    done = nvec * V::N;
  }
  for (int64_t j = done + lane; j < len; j += 32)
    dr[j] = from_f32<T>(fq_elem_bwd<FORM, true>(to_f32<T>(xr[j]), to_f32<T>(gr[j]), p, lo, hi, as, ao));
  as = warp_sum(as);
  ao = warp_sum(ao);
}

}  // namespace dlmcq
