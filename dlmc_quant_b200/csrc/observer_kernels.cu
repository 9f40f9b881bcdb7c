// observer_kernels.cu - PTQ observers (dlmc/quantization/scalar/ops.py) as single-pass kernels.
//
//   obs_stats            one read of x -> {min, max, max|x|, sum|x|} per tensor / per channel
//   *_finalize           O(channels) kernels turning statistics into (scale, offset); multi-GPU
//                        callers all-reduce the statistics between the two calls
//   sweep_tensor         ops.py:36-68   80-candidate clip-ratio MSE search, per tensor: every thread
//                        keeps its elements in registers and evaluates all 80 candidates on them,
//                        so x is read from HBM exactly once (the reference reads it 80 x 9 times)
//   sweep_channel        ops.py:169-196 per-channel search: one warp per row, the row staged once in
//                        shared memory by a bulk async copy (TMA, cp.async.bulk + mbarrier) and swept
//                        80 times on chip, including the reference's sequential accept rule
//   l2norm_step          ops.py:71-83,198-215 one fixed-point iteration (two dot products per row)
//
// Rooflines: stats / l2norm are HBM-bound (4 B/elem fp32).  The sweeps are fp32-issue-bound once
// staged (80 x ~20 instructions per element against 4 B of traffic).
#include "fq_math.cuh"

namespace dlmcq {

constexpr int kNC = DLMCQ_SWEEP_CANDIDATES;

// ---------------------------------------------------------------------------------------
// statistics
// ---------------------------------------------------------------------------------------
struct Stat4 {
  float mn, mx, am, sm;
};
__device__ __forceinline__ void stat_init(Stat4& s) {
  s.mn = INFINITY; s.mx = -INFINITY; s.am = 0.f; s.sm = 0.f;
}
template <bool ABS = false>
__device__ __forceinline__ void stat_add(Stat4& s, float v) {
  const float a = fabsf(v);
  if (ABS) v = a;
  s.mn = fminf(s.mn, v); s.mx = fmaxf(s.mx, v); s.am = fmaxf(s.am, a); s.sm += a;   // NaN reaches sm
}
__device__ __forceinline__ void stat_merge(Stat4& s, const Stat4& o) {
  s.mn = fminf(s.mn, o.mn); s.mx = fmaxf(s.mx, o.mx); s.am = fmaxf(s.am, o.am); s.sm += o.sm;
}
__device__ __forceinline__ Stat4 stat_warp(Stat4 s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Stat4 t;
    t.mn = __shfl_xor_sync(0xffffffffu, s.mn, o);
    t.mx = __shfl_xor_sync(0xffffffffu, s.mx, o);
    t.am = __shfl_xor_sync(0xffffffffu, s.am, o);
    t.sm = __shfl_xor_sync(0xffffffffu, s.sm, o);
    stat_merge(s, t);
  }
  return s;
}
// torch.min / torch.max propagate NaN; fminf/fmaxf do not.  Any NaN input makes sum|x| NaN,
// which is then used to poison the other three statistics.
__device__ __forceinline__ void stat_store(float* out, const Stat4& s) {
  const bool nan = s.sm != s.sm;
  out[0] = nan ? s.sm : s.mn;
  out[1] = nan ? s.sm : s.mx;
  out[2] = nan ? s.sm : s.am;
  out[3] = s.sm;
}

template <typename T, bool ABS>
__global__ void __launch_bounds__(kThreads, 4)
stats_flat_kernel(const T* __restrict__ x, int64_t n, float* __restrict__ stats, void* ws) {
  using V = Vec<T>;
  using raw = typename V::raw;
  __shared__ Stat4 sh[32];
  Stat4 s;
  stat_init(s);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
    const int64_t nvec = n / V::N;
    const raw* xv = reinterpret_cast<const raw*>(x);
    constexpr int U = 4;
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
      raw r[U];
#pragma unroll
      for (int k = 0; k < U; ++k) r[k] = ld_stream(xv + i + k * stride);
#pragma unroll
      for (int k = 0; k < U; ++k) {
        float f[V::N];
        V::unpack(r[k], f);
#pragma unroll
        for (int e = 0; e < V::N; ++e) stat_add<ABS>(s, f[e]);
      }
    }
    for (; i < nvec; i += stride) {
      float f[V::N];
      V::unpack(ld_stream(xv + i), f);
#pragma unroll
      for (int e = 0; e < V::N; ++e) stat_add<ABS>(s, f[e]);
    }
    if (blockIdx.x == 0) {
      const int64_t t = nvec * V::N + threadIdx.x;
      if (t < n) stat_add<ABS>(s, to_f32<T>(x[t]));
    }
  } else {
    for (; i < n; i += stride) stat_add<ABS>(s, to_f32<T>(x[i]));
  }
  s = stat_warp(s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  Stat4* part = reinterpret_cast<Stat4*>(ws_partials(ws));
  if (threadIdx.x == 0) {
    for (int w = 1; w < nwarp; ++w) stat_merge(s, sh[w]);
    part[blockIdx.x] = s;
  }
  if (take_last_ticket(ws_counter(ws), gridDim.x)) {
    Stat4 t;
    stat_init(t);
    for (int b = threadIdx.x; b < static_cast<int>(gridDim.x); b += blockDim.x) stat_merge(t, part[b]);
    t = stat_warp(t);
    __syncthreads();
    if (lane == 0) sh[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < nwarp; ++w) stat_merge(t, sh[w]);
      stat_store(stats, t);
      *ws_counter(ws) = 0u;
    }
  }
}

template <typename T, bool ABS>
__global__ void __launch_bounds__(kRowWarps * 32)
stats_rows_kernel(const T* __restrict__ x, RowGeom gm, float* __restrict__ stats, Stat4* __restrict__ part, int direct) {
  using V = Vec<T>;
  using raw = typename V::raw;
  const int lane = threadIdx.x & 31;
  const int64_t item = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (item >= gm.rows * gm.segs) return;
  const int64_t row = item / gm.segs, seg = item - row * gm.segs;
  const int64_t beg = seg * gm.seg;
  const int64_t len = (gm.inner - beg) < gm.seg ? (gm.inner - beg) : gm.seg;
  const T* xr = x + row * gm.inner + beg;
  Stat4 s;
  stat_init(s);
  int64_t done = 0;
  if ((reinterpret_cast<uintptr_t>(xr) & 15u) == 0) {
    const int64_t nvec = len / V::N;
    const raw* xv = reinterpret_cast<const raw*>(xr);
    for (int64_t j = lane; j < nvec; j += 32) {
      float f[V::N];
      V::unpack(ld_stream(xv + j), f);
#pragma unroll
      for (int e = 0; e < V::N; ++e) stat_add<ABS>(s, f[e]);
    }
    done = nvec * V::N;
  }
  for (int64_t j = done + lane; j < len; j += 32) stat_add<ABS>(s, to_f32<T>(xr[j]));
  s = stat_warp(s);
  if (lane == 0) {
    if (direct) stat_store(stats + 4 * (row % gm.channels), s);
    else part[item] = s;
  }
}

__global__ void __launch_bounds__(kRowWarps * 32)
stats_rows_finalize(const Stat4* __restrict__ part, RowGeom gm, int64_t outer, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t ch = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (ch >= gm.channels) return;
  Stat4 s;
  stat_init(s);
  const int64_t per = outer * gm.segs;
  for (int64_t k = lane; k < per; k += 32) {
    const int64_t b = k / gm.segs, sg = k - b * gm.segs;
    stat_merge(s, part[(b * gm.channels + ch) * gm.segs + sg]);
  }
  s = stat_warp(s);
  if (lane == 0) stat_store(stats + 4 * ch, s);
}

// ops.py:20-34 / 121-140
__global__ void minmax_finalize_kernel(const float* __restrict__ stats, float* __restrict__ scale,
                                       float* __restrict__ offset, int64_t channels, float qdiv, int is_signed,
                                       int allow_offset) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= channels) return;
  const float* s = stats + 4 * c;
  if (is_signed) {
    scale[c] = s[2] / qdiv;            // abs().max() / (2^(n-1)-1)
    offset[c] = 0.f;
  } else {
    const float lo = allow_offset ? s[0] : 0.f;
    scale[c] = (s[1] - lo) / qdiv;     // (max - min) / (2^n - 1)
    offset[c] = lo;
  }
}

// mean|x| * multiplier, computed as (sum / count) then * multiplier like the eager chain
__global__ void absmean_finalize_kernel(const float* __restrict__ stats, float* __restrict__ out, int64_t channels,
                                        float count, float mul_a, float mul_b, int mode) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= channels) return;
  const float mean = stats[4 * c + 3] / count;
  // mode 0: modules/base.py:84   2 * mean / sqrt(qmax)        -> (mul_a * mean) / mul_b
  // mode 1: RootQ/base.py:115    2 * mean * sqrt(qmax)        -> (mul_a * mean) * mul_b
  out[c] = mode == 0 ? (mul_a * mean) / mul_b : (mul_a * mean) * mul_b;
}

// ---------------------------------------------------------------------------------------
// 80-candidate sweep, per tensor
// ---------------------------------------------------------------------------------------
// Candidate i (ops.py:53-58): r = 1 - 0.01*i (python double, cast to fp32 when it meets the
// tensor); c_hi = r*max; c_lo = r*min; scale = (c_hi - c_lo)/qmax; zp = round(-c_lo/scale).
__device__ __forceinline__ void sweep_candidate(int i, float lo, float hi, float qmax, float& sc, float& zp) {
  const float r = static_cast<float>(1.0 - 0.01 * static_cast<double>(i));
  const float c_hi = r * hi, c_lo = r * lo;
  sc = (c_hi - c_lo) / qmax;
  zp = rintf((-c_lo) / sc);
}
// squared error of one element under one candidate (ops.py:59-61), literal chain
__device__ __forceinline__ float sweep_err2(float x, float sc, float zp, float qmax) {
  float q = rintf(x / sc) + zp;
  q = (clamp_ref(q, 0.f, qmax) - zp) * sc;
  const float d = q - x;
  return d * d;
}
// same value on the fast-path domain (|x| <= 2^60, candidate scale in [2^-40, 2^40]): residual-corrected
// division by the candidate's hoisted reciprocal, NaN-propagating min/max clamp (a zero's sign is erased
// by the following "- zp").  11 instructions instead of ~22: the sweep is fp32-issue-bound.
__device__ __forceinline__ float sweep_err2_fast(float x, const FastDiv& fd, float zp, float qmax) {
  float q = rintf(fast_div(x, fd)) + zp;
  q = (clamp_fast(q, 0.f, qmax) - zp) * fd.s;
  const float d = q - x;
  return d * d;
}

constexpr int kSweepElems = 8;   // elements a thread holds in registers per tile (fp32: two 128-bit loads)

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
sweep_tensor_kernel(const T* __restrict__ x, int64_t n, const float* __restrict__ stats, float qmax, int allow_offset,
                    float* __restrict__ sse, void* ws) {
  using V = Vec<T>;
  using raw = typename V::raw;
  constexpr int NV = kSweepElems / V::N >= 1 ? kSweepElems / V::N : 1;   // vectors per thread per tile
  constexpr int NE = NV * V::N;
  __shared__ float c_sc[kNC], c_zp[kNC], c_rc[kNC];
  __shared__ int c_ok[kNC];
  __shared__ float w_acc[kThreads / 32][kNC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < kNC) {
    const float lo = allow_offset ? stats[0] : 0.f;
    sweep_candidate(threadIdx.x, lo, stats[1], qmax, c_sc[threadIdx.x], c_zp[threadIdx.x]);
    const FastDiv fd = make_fastdiv(c_sc[threadIdx.x]);
    c_rc[threadIdx.x] = fd.r;
    c_ok[threadIdx.x] = fd.ok ? 1 : 0;
  }
  for (int k = threadIdx.x; k < (kThreads / 32) * kNC; k += blockDim.x) (&w_acc[0][0])[k] = 0.f;
  __syncthreads();

  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
  const int64_t nvec = vec ? n / V::N : 0;
  const int64_t tiles = nvec / (static_cast<int64_t>(kThreads) * NV);   // full tiles only
  const raw* xv = reinterpret_cast<const raw*>(x);
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    float f[NE];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float tmp[V::N];
      V::unpack(ld_stream(xv + (t * NV + v) * kThreads + threadIdx.x), tmp);
#pragma unroll
      for (int e = 0; e < V::N; ++e) f[v * V::N + e] = tmp[e];
    }
    float m = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e) m = fmaxf(m, fabsf(f[e]));
    const bool tile_ok = m <= kFastDivMaxX;      // inf / huge values: literal chain for this thread's tile
#pragma unroll 2
    for (int c = 0; c < kNC; ++c) {
      const float sc = c_sc[c], zp = c_zp[c];
      float a = 0.f;
      if (tile_ok && c_ok[c]) {
        FastDiv fd;
        fd.s = sc; fd.r = c_rc[c]; fd.ok = true;
#pragma unroll
        for (int e = 0; e < NE; ++e) a += sweep_err2_fast(f[e], fd, zp, qmax);
      } else {
#pragma unroll
        for (int e = 0; e < NE; ++e) a += sweep_err2(f[e], sc, zp, qmax);
      }
      a = warp_sum(a);
      if (lane == 0) w_acc[warp][c] += a;
    }
  }
  // remainder (elements past the last full tile, or everything when x is unaligned): block 0
  if (blockIdx.x == 0) {
    const int64_t start = tiles * kThreads * NE;
    for (int64_t j = start + threadIdx.x; j < start + ((n - start + kThreads - 1) / kThreads) * kThreads; j += kThreads) {
      const bool ok = j < n;
      const float v = ok ? to_f32<T>(x[j]) : 0.f;
      for (int c = 0; c < kNC; ++c) {
        float a = ok ? sweep_err2(v, c_sc[c], c_zp[c], qmax) : 0.f;
        a = warp_sum(a);
        if (lane == 0) w_acc[warp][c] += a;
      }
    }
  }
  __syncthreads();
  float* part = ws_partials(ws);
  if (threadIdx.x < kNC) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) a += w_acc[w][threadIdx.x];
    part[static_cast<int64_t>(blockIdx.x) * kNC + threadIdx.x] = a;
  }
  if (take_last_ticket(ws_counter(ws), gridDim.x)) {
    if (threadIdx.x < kNC) {
      double a = 0.0;
      for (int b = 0; b < static_cast<int>(gridDim.x); ++b) a += part[static_cast<int64_t>(b) * kNC + threadIdx.x];
      sse[threadIdx.x] = static_cast<float>(a);
    }
    if (threadIdx.x == 0) *ws_counter(ws) = 0u;
  }
}

// ops.py:48-66: loss_i = l2_loss = sse_i / rows_for_mean; first strict minimum below 1000 wins,
// otherwise the fallback (max/qmax, 0) of :49-50 stays.
__global__ void sweep_tensor_finalize_kernel(const float* __restrict__ sse, const float* __restrict__ stats,
                                             float rows_for_mean, float qmax, int allow_offset,
                                             float* __restrict__ scale, float* __restrict__ offset,
                                             int32_t* __restrict__ picked) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float lo = allow_offset ? stats[0] : 0.f, hi = stats[1];
  float best = 1000.f, s_out = hi / qmax, z_out = 0.f;
  int pick = -1;
  for (int i = 0; i < kNC; ++i) {
    const float loss = sse[i] / rows_for_mean;
    if (loss < best) {
      best = loss;
      pick = i;
      sweep_candidate(i, lo, hi, qmax, s_out, z_out);
    }
  }
  scale[0] = s_out;
  offset[0] = z_out;
  if (picked) picked[0] = pick;
}

// ---------------------------------------------------------------------------------------
// 80-candidate sweep, per channel: one warp per row, row staged in shared memory by TMA
// ---------------------------------------------------------------------------------------
constexpr int kSweepRowCap = 6144;            // floats of shared memory per warp (24 KB)
constexpr int kSweepWarps = 8;                // 8 x 24 KB = 192 KB of the 227 KB per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <typename T>
__global__ void __launch_bounds__(kSweepWarps * 32, 1)
sweep_channel_kernel(const T* __restrict__ x, int64_t channels, int64_t inner, float qmax, float signed_div,
                     int is_signed, float* __restrict__ scale, float* __restrict__ offset) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ __align__(8) unsigned long long bars[kSweepWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kSweepWarps + warp;
  if (row >= channels) return;                 // whole warp exits together; no block-wide sync below
  float* buf = reinterpret_cast<float*>(dyn_smem) + static_cast<size_t>(warp) * kSweepRowCap;
  const T* xr = x + row * inner;
  const bool staged = inner <= kSweepRowCap - 8;
  const float* src = nullptr;                  // fp32 view of the row: shared-memory copy when it fits
  if (staged) {
    if (sizeof(T) == 4) {
      // 16-byte aligned interior by one bulk async copy; the <=3 head / tail elements by plain loads.
      const uintptr_t a = reinterpret_cast<uintptr_t>(xr);
      const int head = static_cast<int>(((16 - (a & 15u)) & 15u) / 4);       // elements before alignment
      const int64_t h = head < inner ? head : inner;
      const int64_t body = ((inner - h) / 4) * 4;                             // multiple of 16 bytes
      float* dst = buf + 4 - h;                                               // dst+h is 16-byte aligned
      const uint32_t bar = smem_u32(&bars[warp]);
      if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      __syncwarp();
      if (body > 0 && lane == 0) {
        const uint32_t bytes = static_cast<uint32_t>(body * 4);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst + h)), "l"(reinterpret_cast<const float*>(xr) + h), "r"(bytes), "r"(bar)
                     : "memory");
      }
      for (int64_t j = lane; j < h; j += 32) dst[j] = to_f32<T>(xr[j]);
      for (int64_t j = h + body + lane; j < inner; j += 32) dst[j] = to_f32<T>(xr[j]);
      if (body > 0) {
        uint32_t ok = 0;
        while (!ok) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                       "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
        }
      }
      __syncwarp();
      src = dst;
    } else {
      for (int64_t j = lane; j < inner; j += 32) buf[j] = to_f32<T>(xr[j]);   // bf16: convert while staging
      __syncwarp();
      src = buf;
    }
  }
  auto at = [&](int64_t j) -> float { return staged ? src[j] : to_f32<T>(xr[j]); };

  // row statistics (ops.py:171 -> quantize_minmax_channel)
  float mn = INFINITY, mx = -INFINITY, am = 0.f;
  bool nan = false;
  for (int64_t j = lane; j < inner; j += 32) {
    const float v = at(j);
    nan |= (v != v);
    mn = fminf(mn, v); mx = fmaxf(mx, v); am = fmaxf(am, fabsf(v));
  }
  mn = warp_min(mn); mx = warp_max(mx); am = warp_max(am);
  nan = __any_sync(0xffffffffu, nan);
  const bool row_fast = am <= kFastDivMaxX;      // no inf / huge value in the row
  if (nan) mn = mx = am = NAN;
  float cur_scale, cur_off;
  if (is_signed) { cur_scale = am / signed_div; cur_off = 0.f; }           // ops.py:125-127
  else           { cur_off = mn; cur_scale = (mx - mn) / qmax; }           // ops.py:129-136
  float min_v = cur_off;                                                   // ops.py:172 (alias of offset)
  const float max_v = cur_off + cur_scale * qmax;                          // ops.py:173
  float best = 1000.f;                                                     // ops.py:176
  for (int i = 0; i < kNC; ++i) {
    float sc, zp;
    sweep_candidate(i, min_v, max_v, qmax, sc, zp);                        // ops.py:179-185
    float a = 0.f;
    const FastDiv fd = make_fastdiv(sc);
    if (fd.ok && row_fast) {
      for (int64_t j = lane; j < inner; j += 32) a += sweep_err2_fast(at(j), fd, zp, qmax);
    } else {
      for (int64_t j = lane; j < inner; j += 32) a += sweep_err2(at(j), sc, zp, qmax);
    }
    a = warp_sum(a);
    if (best > a) {                                                        // ops.py:191-194
      cur_scale = sc;
      cur_off = zp;
      min_v = zp;           // offset[c] = new_offset also rewrites min_val[c] (reference aliasing)
      best = a;
    }
  }
  if (lane == 0) { scale[row] = cur_scale; offset[row] = cur_off; }
}

// ---------------------------------------------------------------------------------------
// l2norm fixed point: one iteration
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kRowWarps * 32)
l2norm_rows_kernel(const T* __restrict__ x, RowGeom gm, const float* __restrict__ scale,
                   const float* __restrict__ offset, float lo, float hi, const int32_t* __restrict__ done,
                   float* __restrict__ part) {
  using V = Vec<T>;
  using raw = typename V::raw;
  if (*done) return;
  const int lane = threadIdx.x & 31;
  const int64_t item = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (item >= gm.rows * gm.segs) return;
  const int64_t row = item / gm.segs, seg = item - row * gm.segs;
  const ChanParams p = make_params<DLMCQ_FORM_A1>(scale, offset, row % gm.channels, 0.f, lo, hi);
  const int64_t beg = seg * gm.seg;
  const int64_t len = (gm.inner - beg) < gm.seg ? (gm.inner - beg) : gm.seg;
  const T* xr = x + row * gm.inner + beg;
  float a = 0.f, b = 0.f;
  auto add_vec = [&](const float (&f)[V::N]) {
    float code[V::N], y[V::N];
    fq_vec<DLMCQ_FORM_A1, V::N>(f, p, lo, hi, code, y);   // ops.py:78,206 quantize()
#pragma unroll
    for (int e = 0; e < V::N; ++e) {
      a += f[e] * code[e];                                 // (tensor * tensor_q).sum()
      b += code[e] * code[e] + 1e-7f;                      // (tensor_q * tensor_q + 1e-7).sum()
    }
  };
  auto add = [&](float v) {
    float code, y;
    fq_elem<DLMCQ_FORM_A1>(v, p, lo, hi, code, y);
    a += v * code;
    b += code * code + 1e-7f;
  };
  int64_t fin = 0;
  if ((reinterpret_cast<uintptr_t>(xr) & 15u) == 0) {
    const int64_t nvec = len / V::N;
    const raw* xv = reinterpret_cast<const raw*>(xr);
    for (int64_t j = lane; j < nvec; j += 64) {
      const bool two = (j + 32) < nvec;
      raw r0 = ld_stream(xv + j), r1 = r0;
      if (two) r1 = ld_stream(xv + j + 32);
      float f[V::N];
      V::unpack(r0, f);
      add_vec(f);
      if (two) {
        V::unpack(r1, f);
        add_vec(f);
      }
    }
    fin = nvec * V::N;
  }
  for (int64_t j = fin + lane; j < len; j += 32) add(to_f32<T>(xr[j]));
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) { part[2 * item] = a; part[2 * item + 1] = b; }
}

// single CTA: new scale per channel, convergence measure, done flag (ops.py:79-81,207-210).
// One warp per channel with the lanes striding over that channel's segments (per-tensor: the whole
// CTA strides over the segments of the single channel).
__global__ void __launch_bounds__(kThreads)
l2norm_finalize_kernel(const float* __restrict__ part, RowGeom gm, float* __restrict__ scale, float* __restrict__ diff,
                       int32_t* __restrict__ done, int32_t* __restrict__ iters) {
  __shared__ double sh[4][kThreads / 32];
  if (*done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = kThreads / 32;
  double num = 0.0, den = 0.0;
  float single_new = 0.f, single_old = 0.f;
  if (gm.channels == 1) {
    double a = 0.0, b = 0.0;
    for (int64_t sg = threadIdx.x; sg < gm.segs; sg += blockDim.x) {
      a += part[2 * sg];
      b += part[2 * sg + 1];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { sh[2][warp] = a; sh[3][warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double ta = 0.0, tb = 0.0;
      for (int w = 0; w < NW; ++w) { ta += sh[2][w]; tb += sh[3][w]; }
      single_old = scale[0];
      single_new = static_cast<float>(ta) / static_cast<float>(tb);
      scale[0] = single_new;
    }
  } else {
    for (int64_t c = warp; c < gm.channels; c += NW) {
      double a = 0.0, b = 0.0;
      for (int64_t sg = lane; sg < gm.segs; sg += 32) {
        a += part[2 * (c * gm.segs + sg)];
        b += part[2 * (c * gm.segs + sg) + 1];
      }
      a = warp_sum(a);
      b = warp_sum(b);
      if (lane == 0) {
        const float s_old = scale[c];
        const float s_new = static_cast<float>(a) / static_cast<float>(b);
        scale[c] = s_new;
        const float d = s_new - s_old;
        num += static_cast<double>(d * d);
        den += static_cast<double>(s_old * s_old);
      }
    }
    if (lane == 0) { sh[0][warp] = num; sh[1][warp] = den; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float df;
    if (gm.channels == 1) {
      df = fabsf(single_new - single_old) / single_old;                 // ops.py:80
    } else {
      double tn = 0.0, td = 0.0;
      for (int w = 0; w < NW; ++w) { tn += sh[0][w]; td += sh[1][w]; }
      df = sqrtf(static_cast<float>(tn)) / sqrtf(static_cast<float>(td));   // ops.py:209
    }
    diff[0] = df;
    if (iters) iters[0] += 1;
    if (!(df > 1e-5f)) done[0] = 1;                                     // `while diff > epsilon`
  }
}

template <typename T, bool ABS>
static int stats_launch(const T* x, float* stats, const dlmcq_layout* l, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t n = l->outer * l->channels * l->inner;
  if (ws_bytes < dlmcq_workspace_bytes(l)) return DLMCQ_EWORKSPACE;
  if (l->channels == 1) {
    const int64_t tiles = (n / Vec<T>::N + kThreads * 4 - 1) / (kThreads * 4);
    stats_flat_kernel<T, ABS><<<stream_grid(tiles, 8), kThreads, 0, st>>>(x, n, stats, ws);
  } else {
    const RowGeom gm = make_geom(l->outer, l->channels, l->inner);
    const int64_t items = gm.rows * gm.segs;
    const int64_t blocks = (items + kRowWarps - 1) / kRowWarps;
    if (blocks > 0x7fffffffLL) return DLMCQ_EUNSUPPORTED;
    const int direct = (l->outer == 1 && gm.segs == 1) ? 1 : 0;
    Stat4* part = reinterpret_cast<Stat4*>(ws_partials(ws));
    stats_rows_kernel<T, ABS><<<static_cast<unsigned>(blocks), kRowWarps * 32, 0, st>>>(x, gm, stats, part, direct);
    if (!direct) {
      DLMCQ_LAUNCH_CHECK();
      const int64_t fb = (gm.channels + kRowWarps - 1) / kRowWarps;
      stats_rows_finalize<<<static_cast<unsigned>(fb), kRowWarps * 32, 0, st>>>(part, gm, l->outer, stats);
    }
  }
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

}  // namespace dlmcq

using namespace dlmcq;

extern "C" int dlmcq_obs_stats(const void* x, float* stats, const dlmcq_layout* layout, int flags, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (!layout || layout->outer < 1 || layout->channels < 1 || layout->inner < 1) return DLMCQ_EINVAL;
  if (!x || !stats || !workspace) return DLMCQ_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool ab = (flags & DLMCQ_STATS_ABS_INPUT) != 0;
  if (layout->dtype == DLMCQ_F32) {
    const float* p = static_cast<const float*>(x);
    return ab ? stats_launch<float, true>(p, stats, layout, workspace, workspace_bytes, st)
              : stats_launch<float, false>(p, stats, layout, workspace, workspace_bytes, st);
  }
  if (layout->dtype == DLMCQ_BF16) {
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(x);
    return ab ? stats_launch<__nv_bfloat16, true>(p, stats, layout, workspace, workspace_bytes, st)
              : stats_launch<__nv_bfloat16, false>(p, stats, layout, workspace, workspace_bytes, st);
  }
  return DLMCQ_EINVAL;
}

extern "C" int dlmcq_obs_minmax_finalize(const float* stats, float* scale, float* offset, int64_t channels,
                                         int n_bits, int is_signed, int allow_offset, void* stream) {
  if (!stats || !scale || !offset || channels < 1 || n_bits < 1 || n_bits > 24) return DLMCQ_EINVAL;
  const float qdiv = is_signed ? static_cast<float>((1 << (n_bits - 1)) - 1) : static_cast<float>((1 << n_bits) - 1);
  const int blocks = static_cast<int>((channels + 127) / 128);
  minmax_finalize_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(stats, scale, offset, channels, qdiv,
                                                                                is_signed, allow_offset);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

extern "C" int dlmcq_obs_absmean_finalize(const float* stats, float* out, int64_t channels, double count,
                                          double mul_a, double mul_b, int mode, void* stream) {
  if (!stats || !out || channels < 1 || count <= 0) return DLMCQ_EINVAL;
  const int blocks = static_cast<int>((channels + 127) / 128);
  absmean_finalize_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      stats, out, channels, static_cast<float>(count), static_cast<float>(mul_a), static_cast<float>(mul_b), mode);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

extern "C" int dlmcq_obs_sweep_tensor_sse(const void* x, int64_t numel, int dtype, const float* stats, int n_bits,
                                          int allow_offset, float* sse, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  if (!x || !stats || !sse || !workspace || numel < 1 || n_bits < 1 || n_bits > 24) return DLMCQ_EINVAL;
  if (workspace_bytes < dlmcq_workspace_bytes(nullptr)) return DLMCQ_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float qmax = static_cast<float>((1 << n_bits) - 1);
  const int64_t tiles = (numel + kThreads * kSweepElems - 1) / (kThreads * kSweepElems);
  const int grid = stream_grid(tiles, 2);
  if (dtype == DLMCQ_F32)
    sweep_tensor_kernel<float><<<grid, kThreads, 0, st>>>(static_cast<const float*>(x), numel, stats, qmax,
                                                          allow_offset, sse, workspace);
  else if (dtype == DLMCQ_BF16)
    sweep_tensor_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x), numel, stats,
                                                                  qmax, allow_offset, sse, workspace);
  else
    return DLMCQ_EINVAL;
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

extern "C" int dlmcq_obs_sweep_tensor_finalize(const float* sse, const float* stats, double rows_for_mean, int n_bits,
                                               int allow_offset, float* scale, float* offset, int32_t* picked,
                                               void* stream) {
  if (!sse || !stats || !scale || !offset || rows_for_mean <= 0 || n_bits < 1 || n_bits > 24) return DLMCQ_EINVAL;
  sweep_tensor_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      sse, stats, static_cast<float>(rows_for_mean), static_cast<float>((1 << n_bits) - 1), allow_offset, scale,
      offset, picked);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

extern "C" int dlmcq_obs_sweep_channel(const void* x, int64_t channels, int64_t inner, int dtype, int n_bits,
                                       int is_signed, float* scale, float* offset, void* stream) {
  if (!x || !scale || !offset || channels < 1 || inner < 1 || n_bits < 1 || n_bits > 24) return DLMCQ_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float qmax = static_cast<float>((1 << n_bits) - 1);
  const float sdiv = static_cast<float>((1 << (n_bits - 1)) - 1);
  const size_t smem = static_cast<size_t>(kSweepWarps) * kSweepRowCap * sizeof(float);
  const int64_t blocks = (channels + kSweepWarps - 1) / kSweepWarps;
  if (blocks > 0x7fffffffLL) return DLMCQ_EUNSUPPORTED;
  cudaError_t e;
  if (dtype == DLMCQ_F32) {
    e = cudaFuncSetAttribute(sweep_channel_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e);
    sweep_channel_kernel<float><<<static_cast<unsigned>(blocks), kSweepWarps * 32, smem, st>>>(
        static_cast<const float*>(x), channels, inner, qmax, sdiv, is_signed, scale, offset);
  } else if (dtype == DLMCQ_BF16) {
    e = cudaFuncSetAttribute(sweep_channel_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e);
    sweep_channel_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), kSweepWarps * 32, smem, st>>>(
        static_cast<const __nv_bfloat16*>(x), channels, inner, qmax, sdiv, is_signed, scale, offset);
  } else {
    return DLMCQ_EINVAL;
  }
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

extern "C" int dlmcq_obs_l2norm_step(const void* x, int64_t channels, int64_t inner, int dtype, float* scale,
                                     const float* offset, int lo, int hi, float* diff, int32_t* done, int32_t* iters,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !scale || !diff || !done || !workspace || channels < 1 || inner < 1) return DLMCQ_EINVAL;
  dlmcq_layout l = {1, channels, inner, dtype};
  if (workspace_bytes < dlmcq_workspace_bytes(&l)) return DLMCQ_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const RowGeom gm = make_geom(1, channels, inner);
  const int64_t items = gm.rows * gm.segs;
  const int64_t blocks = (items + kRowWarps - 1) / kRowWarps;
  if (blocks > 0x7fffffffLL) return DLMCQ_EUNSUPPORTED;
  float* part = ws_partials(workspace);
  const float flo = static_cast<float>(lo), fhi = static_cast<float>(hi);
  if (dtype == DLMCQ_F32)
    l2norm_rows_kernel<float><<<static_cast<unsigned>(blocks), kRowWarps * 32, 0, st>>>(
        static_cast<const float*>(x), gm, scale, offset, flo, fhi, done, part);
  else if (dtype == DLMCQ_BF16)
    l2norm_rows_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), kRowWarps * 32, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), gm, scale, offset, flo, fhi, done, part);
  else
    return DLMCQ_EINVAL;
  DLMCQ_LAUNCH_CHECK();
  l2norm_finalize_kernel<<<1, kThreads, 0, st>>>(part, gm, scale, diff, done, iters);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}
