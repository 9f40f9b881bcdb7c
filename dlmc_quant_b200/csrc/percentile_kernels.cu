// percentile_kernels.cu - exact order statistics (k-th smallest value) of a tensor by a 3-pass radix select,
// the building block of the percentile-clipping observer.
//
// The north star lists a percentile observer next to min/max and the MSE sweep; the reference itself
// (dlmc/quantization/scalar/ops.py) has none, so the semantics are defined here and pinned against
// torch.kthvalue: value_j = the ranks[j]-th smallest element (1-based) of x, or of |x|.  Exact - no sampling, no
// histogram interpolation - so that the qparams derived from it are reproducible bit for bit.
//
//   key(x)    order-preserving uint32 image of the fp32 value (NaN last, like torch.sort; -0 == +0)
//   pass p    histogram of key digit p (11, 11, 10 bits from the top) over the elements whose higher digits
//             equal the prefix selected so far; one read of x per pass (3 x 4 B/elem fp32, 3 x 2 B/elem bf16)
//   select    the digit whose bin holds the wanted rank; prefix <- prefix.digit, rank <- rank - (elements below)
// Up to two ranks are tracked at once (lower and upper percentile share the passes).  Multi-GPU callers
// all-reduce (SUM) the histogram between `hist` and `select`, which yields the order statistic of the union.
// Roofline: HBM (three streaming reads); the shared-memory histogram is the cost: real activations put half of
// their elements into one bin (post-ReLU zeros, counted in registers) and the rest into a few dozen bins.
#include "common.cuh"
#include "fq_math.cuh"

namespace dlmcq {

constexpr int kRadixBins = 2048;
constexpr int kRadixRep = 8;
__host__ __device__ inline int radix_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
__host__ __device__ inline unsigned radix_mask(int pass) { return pass == 2 ? 1023u : 2047u; }

struct RadixState {
  unsigned long long rank[2];   // remaining 1-based rank inside the selected prefix
  unsigned int prefix[2];       // selected high digits, right-aligned
  unsigned int n_ranks;
  unsigned int pad;
};

__device__ __forceinline__ unsigned int radix_key(float v, bool abs_input) {
  if (abs_input) v = fabsf(v);
  if (v != v) return 0xffffffffu;                  // NaN sorts last
  if (v == 0.f) v = 0.f;                           // -0 and +0 are the same value
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float radix_unkey(unsigned int k) {
  if (k == 0xffffffffu) return __uint_as_float(0x7fc00000u);
  const unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

constexpr unsigned int kZeroKey = 0x80000000u;   // key(+-0.0): the one value real activations repeat (post-ReLU)

template <typename T>
__global__ void __launch_bounds__(kThreads, 3)
radix_hist_kernel(const T* __restrict__ x, int64_t n, int abs_input, int pass, const RadixState* __restrict__ state,
                  unsigned int* __restrict__ hist /* [2][kRadixBins] */) {
  using V = Vec<T>;
  using raw = typename V::raw;
  // kRadixRep copies of the first rank's histogram (copy = lane & 7): a warp's lanes hit the same few dozen
  // bins (the digit is sign + exponent + two mantissa bits), the copies cut the same-address conflicts by 8
  extern __shared__ unsigned int sh_dyn[];
  unsigned int* sh0 = sh_dyn + (threadIdx.x & (kRadixRep - 1)) * kRadixBins;
  unsigned int* sh1 = sh_dyn + kRadixRep * kRadixBins;
  for (int k = threadIdx.x; k < (kRadixRep + 1) * kRadixBins; k += blockDim.x) sh_dyn[k] = 0u;
  __syncthreads();
  const RadixState st = *state;
  const int shift = radix_shift(pass);
  const unsigned int mask = radix_mask(pass);
  const int up = pass == 0 ? 32 : radix_shift(pass - 1);      // bits above this digit belong to the prefix
  const bool two = st.n_ranks > 1 && pass > 0 && st.prefix[0] != st.prefix[1];
  // Zeros are counted in a register and added once per thread; everything else goes through plain
  // shared-memory atomics (measured 3x faster than warp-aggregated match.any atomics on random data).
  unsigned int z0 = 0, z1 = 0;
  auto add = [&](float v, bool ok) {
    const unsigned int key = radix_key(v, abs_input != 0);
    const unsigned int hi = pass == 0 ? 0u : (key >> up);
    const bool a0 = ok && (pass == 0 || hi == st.prefix[0]);
    const bool a1 = two && ok && hi == st.prefix[1];
    if (key == kZeroKey) {
      z0 += a0 ? 1u : 0u;
      z1 += a1 ? 1u : 0u;
    } else {
      const unsigned int bin = (key >> shift) & mask;
      if (a0) atomicAdd(sh0 + bin, 1u);
      if (a1) atomicAdd(sh1 + bin, 1u);
    }
  };
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
    const int64_t nvec = n / V::N;
    const raw* xv = reinterpret_cast<const raw*>(x);
    constexpr int U = 4;
    const int64_t full = nvec / (U * stride) * (U * stride);   // every thread of the grid takes part in these
    for (; i < full; i += U * stride) {
      raw r[U];
#pragma unroll
      for (int k = 0; k < U; ++k) r[k] = ld_stream(xv + i + k * stride);
#pragma unroll
      for (int k = 0; k < U; ++k) {
        float f[V::N];
        V::unpack(r[k], f);
#pragma unroll
        for (int e = 0; e < V::N; ++e) add(f[e], true);
      }
    }
    // tail
    const int64_t tail_iters = (nvec - full + stride - 1) / stride;
    for (int64_t t = 0; t < tail_iters; ++t, i += stride) {
      const bool ok = i < nvec;
      float f[V::N];
      if (ok) V::unpack(ld_stream(xv + i), f);
#pragma unroll
      for (int e = 0; e < V::N; ++e) add(ok ? f[e] : 0.f, ok);
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
      const int64_t t = nvec * V::N + threadIdx.x;
      add(t < n ? to_f32<T>(x[t]) : 0.f, t < n);
    }
  } else {
    const int64_t iters = (n + stride - 1) / stride;
    for (int64_t t = 0; t < iters; ++t, i += stride) add(i < n ? to_f32<T>(x[i]) : 0.f, i < n);
  }
  const unsigned int zbin = (kZeroKey >> shift) & mask;
  if (z0) atomicAdd(sh0 + zbin, z0);
  if (z1) atomicAdd(sh1 + zbin, z1);
  __syncthreads();
  for (int k = threadIdx.x; k < kRadixBins; k += blockDim.x) {
    unsigned int c = 0;
#pragma unroll
    for (int r = 0; r < kRadixRep; ++r) c += sh_dyn[r * kRadixBins + k];
    if (c) atomicAdd(hist + k, c);
    const unsigned int c1 = sh_dyn[kRadixRep * kRadixBins + k];
    if (c1) atomicAdd(hist + kRadixBins + k, c1);
  }
}

// one warp per rank: locate the bin that holds the rank, descend into it
__global__ void __launch_bounds__(64)
radix_select_kernel(const unsigned int* __restrict__ hist, int pass, RadixState* __restrict__ state) {
  const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned int p0 = state->prefix[0], p1 = state->prefix[1], nr = state->n_ranks;
  __syncthreads();                                             // both warps have read the prefixes before either writes
  if (j >= static_cast<int>(nr)) return;
  const bool own = j == 1 && pass > 0 && p0 != p1;
  const unsigned int* h = hist + (own ? kRadixBins : 0);
  const int bins = static_cast<int>(radix_mask(pass)) + 1;
  const int per = bins / 32;
  unsigned long long mine = 0;
  for (int b = 0; b < per; ++b) mine += h[lane * per + b];
  unsigned long long incl = mine;                              // inclusive prefix over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const unsigned long long rank = state->rank[j];
  const unsigned long long before = incl - mine;
  const bool here = rank > before && rank <= incl;             // exactly one lane (rank <= total by construction)
  const unsigned int who = __ballot_sync(0xffffffffu, here);
  if (who == 0) {                                              // rank beyond the population: clamp to the last bin
    if (lane == 31) { state->prefix[j] = (pass == 0 ? 0u : state->prefix[j] << (pass == 2 ? 10 : 11)) | (bins - 1); }
    return;
  }
  if (here) {
    unsigned long long cum = before;
    int b = 0;
    for (; b < per; ++b) {
      const unsigned int c = h[lane * per + b];
      if (rank <= cum + c) break;
      cum += c;
    }
    if (b == per) b = per - 1;
    const unsigned int digit = static_cast<unsigned int>(lane * per + b);
    state->prefix[j] = (pass == 0 ? 0u : state->prefix[j] << (pass == 2 ? 10 : 11)) | digit;
    state->rank[j] = rank - cum;
  }
}

__global__ void radix_init_kernel(RadixState* state, unsigned long long r0, unsigned long long r1, unsigned int n) {
  state->rank[0] = r0; state->rank[1] = r1; state->prefix[0] = 0u; state->prefix[1] = 0u; state->n_ranks = n; state->pad = 0u;
}
__global__ void radix_values_kernel(const RadixState* __restrict__ state, float* __restrict__ values) {
  if (threadIdx.x < state->n_ranks) values[threadIdx.x] = radix_unkey(state->prefix[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------
// Fast path: ONE full read instead of three.
//
// An order statistic near the tails (what a percentile-clipping observer asks for) can be bracketed from a small
// sample: with S pseudo-randomly placed samples the k-th of N elements lies, with probability ~1 - 1e-9, between the
// sample order statistics of rank k*S/N -+ (6*sqrt(S*p*(1-p)) + 3).  So:
//   1. kth_sample_kernel   (1 CTA): gather S = 16384 keys, radix-select the two bracketing sample keys per rank;
//   2. kth_collect_kernel  (full read, streaming): per rank count the elements BELOW the bracket (registers) and append
//                          the few inside it (<= ~1 % of N for p >= 0.99; warp-aggregated appends);
//   3. kth_resolve_kernel  (1 CTA): the answer is the (k - below)-th smallest candidate - exact radix select over the
//                          candidate keys; if the bracket missed (k - below outside [1, #candidates], or the
//                          candidate buffer overflowed) the status word says so and the caller runs the three-pass
//                          histogram select above.  Exactness never depends on the sample, only the speed does.
// ---------------------------------------------------------------------------------------
constexpr int kKthSample = 16384;
constexpr int kKthCtaThreads = 1024;

struct KthFastState {
  unsigned int lo_key[2], hi_key[2];     // bracket per rank (inclusive)
  unsigned long long below[2];           // elements with key < lo_key
  unsigned long long eq_lo[2], eq_hi[2]; // elements equal to a bracket end: counted, never stored (post-ReLU zeros,
                                         // saturated maxima: one value may be half of the tensor)
  unsigned int n_cand[2];                // elements strictly inside the bracket (stored)
  unsigned int overflow;
  unsigned int n_ranks;
  unsigned long long rank[2];
};

// Digit selection shared by the single-CTA and the cooperative selects: `hist` (2048 bins in shared memory, unused
// bins zero) holds the histogram of digit `pass` among the keys that match `prefix`; finds the bin that contains the
// 1-based rank *rank_io (shared), appends its digit to `prefix` and reduces the rank.  blockDim.x == 1024.
__device__ unsigned int cta_scan_select(const unsigned int* hist, unsigned long long* rank_io, unsigned int prefix,
                                        int pass) {
  __shared__ unsigned int s_warp[32];
  __shared__ unsigned int s_digit;
  const unsigned int b0 = hist[2 * threadIdx.x], b1 = hist[2 * threadIdx.x + 1];
  unsigned int incl = b0 + b1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;                           // inclusive over warps
  }
  __syncthreads();
  const unsigned long long rank = *rank_io;
  const unsigned long long excl = (warp ? s_warp[warp - 1] : 0u) + static_cast<unsigned long long>(incl - (b0 + b1));
  __syncthreads();
  if (rank > excl && rank <= excl + b0 + b1) {  // exactly one thread (rank <= population by construction)
    const bool second = rank > excl + b0;
    s_digit = 2 * threadIdx.x + (second ? 1 : 0);
    *rank_io = rank - excl - (second ? b0 : 0u);
  }
  __syncthreads();
  const unsigned int out = (pass == 0 ? 0u : prefix << (pass == 2 ? 10 : 11)) | s_digit;
  __syncthreads();
  return out;
}

// r-th smallest (1-based) of keys[0..n) by an 11/11/10-bit radix select run by one whole CTA (blockDim 1024);
// hist: 2048 unsigned ints of shared memory.  Every thread returns the key.
__device__ unsigned int cta_radix_select(const unsigned int* __restrict__ keys, unsigned int n, unsigned long long r,
                                         unsigned int* hist) {
  __shared__ unsigned long long s_rank;
  unsigned int prefix = 0;
  if (threadIdx.x == 0) s_rank = r;
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = radix_shift(pass);
    const unsigned int mask = radix_mask(pass);
    const int up = pass == 0 ? 32 : radix_shift(pass - 1);
    for (int k = threadIdx.x; k < kRadixBins; k += blockDim.x) hist[k] = 0u;
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned int key = keys[i];
      if (pass == 0 || (key >> up) == prefix) atomicAdd(hist + ((key >> shift) & mask), 1u);
    }
    __syncthreads();
    prefix = cta_scan_select(hist, &s_rank, prefix, pass);
  }
  return prefix;
}

template <typename T>
__global__ void __launch_bounds__(kKthCtaThreads)
kth_sample_kernel(const T* __restrict__ x, int64_t n, int abs_input, unsigned long long rank0, unsigned long long rank1,
                  unsigned int n_ranks, KthFastState* __restrict__ st, unsigned int* __restrict__ sample) {
  __shared__ unsigned int hist[kRadixBins];
  for (int i = threadIdx.x; i < kKthSample; i += blockDim.x) {
    // multiplicative hash -> a fixed pseudo-random position in [0, n): regular strides would alias with periodic data
    const unsigned int hsh = static_cast<unsigned int>(i) * 2654435761u + 0x9e3779b9u;
    const int64_t idx = static_cast<int64_t>((static_cast<unsigned long long>(hsh) * static_cast<unsigned long long>(n)) >> 32);
    sample[i] = radix_key(to_f32<T>(x[idx]), abs_input != 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->below[0] = st->below[1] = 0ull;
    st->eq_lo[0] = st->eq_lo[1] = st->eq_hi[0] = st->eq_hi[1] = 0ull;
    st->n_cand[0] = st->n_cand[1] = 0u;
    st->overflow = 0u;
    st->n_ranks = n_ranks;
    st->rank[0] = rank0;
    st->rank[1] = rank1;
  }
  for (unsigned int j = 0; j < n_ranks; ++j) {
    const double k = static_cast<double>(j == 0 ? rank0 : rank1);
    const double p = k / static_cast<double>(n);
    const double ks = p * kKthSample;
    const double delta = 6.0 * sqrt(kKthSample * p * (1.0 - p)) + 3.0;
    const long long i_lo = static_cast<long long>(floor(ks - delta)), i_hi = static_cast<long long>(ceil(ks + delta));
    unsigned int lo_key = 0u, hi_key = 0xffffffffu;
    if (i_lo >= 1) lo_key = cta_radix_select(sample, kKthSample, static_cast<unsigned long long>(i_lo), hist);
    if (i_hi <= kKthSample) hi_key = cta_radix_select(sample, kKthSample, static_cast<unsigned long long>(i_hi), hist);
    if (threadIdx.x == 0) { st->lo_key[j] = lo_key; st->hi_key[j] = hi_key; }
  }
}

// One rank's share of a vector: counters from float compares (the key order is the float order; NaN fails every
// compare and is - correctly - neither below nor inside any bracket), and - rarely - the append of the elements strictly
// inside the bracket: per-lane count -> warp exclusive scan -> ONE atomic per warp -> ordered stores.
struct KthRank {
  unsigned int lo, hi;       // bracket keys
  float flo, fhi;            // their float images
  unsigned int below, el, eh;
};
// Candidates (keys strictly inside the bracket: a fraction of a percent of the tensor) are appended to a per-CTA
// shared-memory list with a shared-memory atomic - no warp-wide step, so a lane that finds one does not drag its warp
// through a scan and a global atomic's round trip (with a 0.3 % bracket a third of all 128-element warp visits hold
// at least one) - and the list is flushed once, coalesced, when the CTA is done.  Order is irrelevant: the resolve step
// histograms the keys.
constexpr unsigned int kCandSmem = 3072;            // keys per rank per CTA before appends go to global memory directly
// Appending one candidate - the rare case (a fraction of a percent of the elements), kept out of line so that the
// streaming loop stays small (it was missing the instruction cache) and takes its operand in a register.
template <bool ABS>
__device__ __noinline__ void kth_append(float v, unsigned int lo, unsigned int hi, unsigned int* s_cnt,
                                        unsigned int* s_buf, unsigned int* __restrict__ n_cand,
                                        unsigned int* __restrict__ cand, unsigned int cap,
                                        unsigned int* __restrict__ overflow) {
  const unsigned int key = radix_key(v, ABS);
  if (key > lo && key < hi) {
    const unsigned int pos = atomicAdd(s_cnt, 1u);
    if (pos < kCandSmem) {
      s_buf[pos] = key;
    } else {
      const unsigned int gp = atomicAdd(n_cand, 1u);
      if (gp < cap) cand[gp] = key; else *overflow = 1u;
    }
  }
}
// One vector against one bracket.  The common cases are decided from the vector's minimum and maximum: everything
// above the bracket (nothing to count) or everything below it (one add).  fminf ignores a NaN (a NaN sorts last:
// "above" is right for it); the maximum propagates it, so a vector holding one is never taken for "all below".
// Otherwise four compares per element give below / equal-lo / equal-hi (mass points such as post-ReLU zeros sit on
// a bracket end and are only counted), and only a vector with something strictly inside calls kth_append.
template <int N, bool ABS>
__device__ __forceinline__ void kth_visit_vec(const float (&f)[N], float vmin, float vmax, KthRank& r,
                                              unsigned int* s_cnt, unsigned int* s_buf,
                                              unsigned int* __restrict__ n_cand, unsigned int* __restrict__ cand,
                                              unsigned int cap, unsigned int* __restrict__ overflow) {
  if (vmin > r.fhi) return;
  if (vmax < r.flo) { r.below += N; return; }
  unsigned int nb = 0, nle = 0, nlh = 0, nleh = 0;
#pragma unroll
  for (int e = 0; e < N; ++e) {
    const float v = ABS ? fabsf(f[e]) : f[e];
    nb += v < r.flo ? 1u : 0u;
    nle += v <= r.flo ? 1u : 0u;
    nlh += v < r.fhi ? 1u : 0u;
    nleh += v <= r.fhi ? 1u : 0u;
  }
  r.below += nb;
  r.el += nle - nb;
  if (r.hi != r.lo) r.eh += nleh - nlh;          // a degenerate bracket: equality is counted once
  if (nlh > nle || r.fhi != r.fhi) {             // something strictly inside (or a NaN bracket end: decide by key)
#pragma unroll
    for (int e = 0; e < N; ++e) {
      const float v = ABS ? fabsf(f[e]) : f[e];
      if (!(v <= r.flo) && !(v >= r.fhi))          // NaN ends / NaN values pass: kth_append decides by key
        kth_append<ABS>(f[e], r.lo, r.hi, s_cnt, s_buf, n_cand, cand, cap, overflow);
    }
  }
}

template <typename T, bool ABS, bool TWO>
__global__ void __launch_bounds__(kThreads, 4)
kth_collect_kernel(const T* __restrict__ x, int64_t n, KthFastState* __restrict__ st,
                   unsigned int* __restrict__ cand /* [2][cap] */, unsigned int cap) {
  using V = Vec<T>;
  using raw = typename V::raw;
  const int lane = threadIdx.x & 31;
  __shared__ unsigned int s_cnt[2], s_base[2];
  __shared__ unsigned int s_buf[TWO ? 2 : 1][kCandSmem];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0u;
  __syncthreads();
  KthRank r0{st->lo_key[0], st->hi_key[0], radix_unkey(st->lo_key[0]), radix_unkey(st->hi_key[0]), 0u, 0u, 0u};
  KthRank r1{0u, 0u, 0.f, 0.f, 0u, 0u, 0u};
  if (TWO) r1 = KthRank{st->lo_key[1], st->hi_key[1], radix_unkey(st->lo_key[1]), radix_unkey(st->hi_key[1]), 0u, 0u, 0u};
  auto visit = [&](const float (&f)[V::N], bool ok) {
    if (!ok) return;
    float vmin = ABS ? fabsf(f[0]) : f[0], vmax = vmin;
#pragma unroll
    for (int e = 1; e < V::N; ++e) {
      const float v = ABS ? fabsf(f[e]) : f[e];
      vmin = fminf(vmin, v);
      vmax = max_nan(vmax, v);
    }
    kth_visit_vec<V::N, ABS>(f, vmin, vmax, r0, &s_cnt[0], s_buf[0], &st->n_cand[0], cand, cap, &st->overflow);
    if (TWO) kth_visit_vec<V::N, ABS>(f, vmin, vmax, r1, &s_cnt[1], s_buf[TWO ? 1 : 0], &st->n_cand[1], cand + cap, cap,
                                      &st->overflow);
  };
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
  const int64_t nvec = aligned ? n / V::N : 0;
  const raw* xv = reinterpret_cast<const raw*>(x);
  constexpr int U = 4;
  const int64_t full = nvec / (U * stride) * (U * stride);       // whole warps take part in every shuffle
  for (; i < full; i += U * stride) {
    __syncwarp();            // lanes that took the rare paths rejoin here: without it the warp stays split into
                             // sub-warps that run the rest of the loop separately (measured: 8 active lanes on average)
    raw r[U];
#pragma unroll
    for (int k = 0; k < U; ++k) r[k] = ld_stream(xv + i + k * stride);
#pragma unroll
    for (int k = 0; k < U; ++k) {
      float f[V::N];
      V::unpack(r[k], f);
      visit(f, true);
      __syncwarp();
    }
  }
  const int64_t tail_iters = (nvec - full + stride - 1) / stride;
  for (int64_t t = 0; t < tail_iters; ++t, i += stride) {
    __syncwarp();
    const bool ok = i < nvec;
    float f[V::N];
#pragma unroll
    for (int e = 0; e < V::N; ++e) f[e] = 0.f;
    if (ok) V::unpack(ld_stream(xv + i), f);
    visit(f, ok);
  }
  // scalar remainder (and everything, for an unaligned pointer): one element per lane in slot 0 of a vector
  const int64_t rest0 = nvec * V::N;
  const int64_t rest_iters = (n - rest0 + stride - 1) / stride;
  int64_t j = rest0 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (int64_t t = 0; t < rest_iters; ++t, j += stride) {
    const bool ok = j < n;
    float f[V::N];
    // padding values must classify as "above" so that they change no counter: NaN does (fails every compare)
    const float pad = __uint_as_float(0x7fc00000u);
#pragma unroll
    for (int e = 0; e < V::N; ++e) f[e] = pad;
    if (ok) f[0] = to_f32<T>(x[j]);
    // count only slot 0: temporarily classify the padded vector, then undo nothing (NaN pads add to no counter)
    visit(f, ok);
  }
  // flush the CTA's candidate lists: one global atomic per rank, coalesced copies
  __syncthreads();
  if (threadIdx.x < (TWO ? 2 : 1)) {
    const unsigned int c = s_cnt[threadIdx.x] < kCandSmem ? s_cnt[threadIdx.x] : kCandSmem;
    s_cnt[threadIdx.x] = c;
    s_base[threadIdx.x] = c ? atomicAdd(&st->n_cand[threadIdx.x], c) : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < (TWO ? 2 : 1); ++q) {
    const unsigned int c = s_cnt[q], base = s_base[q];
    for (unsigned int t = threadIdx.x; t < c; t += blockDim.x) {
      if (base + t < cap) cand[static_cast<size_t>(q) * cap + base + t] = s_buf[q][t]; else st->overflow = 1u;
    }
  }
  unsigned long long c[6] = {r0.below, r1.below, r0.el, r1.el, r0.eh, r1.eh};
#pragma unroll
  for (int q = 0; q < 6; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c[q] += __shfl_xor_sync(0xffffffffu, c[q], o);
  }
  if (lane == 0) {
    if (c[0]) atomicAdd(&st->below[0], c[0]);
    if (c[1]) atomicAdd(&st->below[1], c[1]);
    if (c[2]) atomicAdd(&st->eq_lo[0], c[2]);
    if (c[3]) atomicAdd(&st->eq_lo[1], c[3]);
    if (c[4]) atomicAdd(&st->eq_hi[0], c[4]);
    if (c[5]) atomicAdd(&st->eq_hi[1], c[5]);
  }
}

// Resolve, cooperatively: every CTA histograms its slice of the keys (shared memory, then the non-empty bins into a
// global histogram), a grid barrier, and every CTA finds the same digit from the same global histogram - three digits
// per rank, six barriers at most, whatever the key count (a single CTA needed ~100 us per pass for the ~0.3 % of a
// 2^28-element tensor a tail bracket collects).  The keys are the collected candidates when the bracket held; when it
// missed (adversarially ordered data, a candidate overflow) the SAME three-digit select runs over the whole tensor
// inside this launch - the values are exact either way and no second launch sequence or host read is needed; the
// status word only reports which of the two happened.
template <typename KeyAt>
__device__ unsigned int grid_radix_select(KeyAt key_at, unsigned long long count, unsigned long long r,
                                          unsigned int* __restrict__ gh3 /* [3][2048], zeroed */,
                                          unsigned int* hist, unsigned long long* s_rank, unsigned int* counter,
                                          unsigned int& barrier_no) {
  unsigned int prefix = 0;
  if (threadIdx.x == 0) *s_rank = r;
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = radix_shift(pass);
    const unsigned int mask = radix_mask(pass);
    const int up = pass == 0 ? 32 : radix_shift(pass - 1);
    unsigned int* gh = gh3 + static_cast<size_t>(pass) * kRadixBins;
    for (int b = threadIdx.x; b < kRadixBins; b += blockDim.x) hist[b] = 0u;
    __syncthreads();
    unsigned int zeros = 0;                                // the one value real data repeats: counted in a register
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
      const unsigned int kk = key_at(i);
      if (pass == 0 || (kk >> up) == prefix) {
        if (kk == kZeroKey) ++zeros;
        else atomicAdd(hist + ((kk >> shift) & mask), 1u);
      }
    }
    if (zeros) atomicAdd(hist + ((kZeroKey >> shift) & mask), zeros);
    __syncthreads();
    for (int b = threadIdx.x; b < kRadixBins; b += blockDim.x)
      if (hist[b]) atomicAdd(gh + b, hist[b]);
    barrier_no += 1;
    grid_barrier(counter, barrier_no * gridDim.x);
    for (int b = threadIdx.x; b < kRadixBins; b += blockDim.x) hist[b] = __ldcg(gh + b);
    __syncthreads();
    prefix = cta_scan_select(hist, s_rank, prefix, pass);
  }
  return prefix;
}

template <typename T, bool ABS>
__global__ void __launch_bounds__(kKthCtaThreads, 1)
kth_resolve_kernel(const KthFastState* __restrict__ st, const unsigned int* __restrict__ cand, unsigned int cap,
                   unsigned int* __restrict__ ghist /* [2 ranks][3 passes][2048], zeroed */,
                   unsigned int* __restrict__ counter /* zeroed */, const T* __restrict__ x, int64_t n,
                   float* __restrict__ values, int32_t* __restrict__ status) {
  __shared__ unsigned int hist[kRadixBins];
  __shared__ unsigned long long s_rank;
  const bool no_overflow = st->overflow == 0u;
  bool all_held = true;
  const unsigned int nr = st->n_ranks;
  unsigned int barrier_no = 0;
  for (unsigned int j = 0; j < nr; ++j) {                  // every branch below is uniform across the grid
    const unsigned long long k = st->rank[j], below = st->below[j], el = st->eq_lo[j], eh = st->eq_hi[j];
    const unsigned int nc = st->n_cand[j];
    unsigned int* gh3 = ghist + static_cast<size_t>(j) * 3 * kRadixBins;
    unsigned int key;
    if (!no_overflow || nc > cap || k <= below || k > below + el + nc + eh) {   // the bracket missed
      all_held = false;
      key = grid_radix_select([&](unsigned long long i) { return radix_key(to_f32<T>(x[i]), ABS); },
                              static_cast<unsigned long long>(n), k, gh3, hist, &s_rank, counter, barrier_no);
    } else {
      const unsigned long long r = k - below;
      if (r <= el) {
        key = st->lo_key[j];
      } else if (r > el + nc) {
        key = st->hi_key[j];
      } else {
        const unsigned int* keys = cand + static_cast<size_t>(j) * cap;
        key = grid_radix_select([&](unsigned long long i) { return __ldcg(keys + i); }, nc, r - el, gh3, hist, &s_rank,
                                counter, barrier_no);
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) values[j] = radix_unkey(key);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) status[0] = all_held ? 1 : 0;
}

}  // namespace dlmcq

using namespace dlmcq;

extern "C" size_t dlmcq_obs_kth_state_bytes(void) { return 256 + 2 * kRadixBins * sizeof(unsigned int); }

extern "C" int dlmcq_obs_kth_begin(void* state, int64_t rank0, int64_t rank1, void* stream) {
  if (!state || rank0 < 1 || rank1 < 0) return DLMCQ_EINVAL;
  radix_init_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<RadixState*>(state), static_cast<unsigned long long>(rank0), static_cast<unsigned long long>(rank1),
      rank1 > 0 ? 2u : 1u);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

extern "C" int dlmcq_obs_kth_hist(const void* x, int64_t numel, int dtype, int flags, int pass, void* state,
                                  void* stream) {
  if (!x || !state || numel < 1 || pass < 0 || pass > 2) return DLMCQ_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned int* hist = reinterpret_cast<unsigned int*>(static_cast<char*>(state) + 256);
  cudaError_t e = cudaMemsetAsync(hist, 0, 2 * kRadixBins * sizeof(unsigned int), st);
  if (e != cudaSuccess) return set_cuda_error(e);
  const int abs_input = (flags & DLMCQ_STATS_ABS_INPUT) ? 1 : 0;
  const RadixState* rs = static_cast<const RadixState*>(state);
  const size_t smem = static_cast<size_t>(kRadixRep + 1) * kRadixBins * sizeof(unsigned int);   // 72 KB
  if (dtype == DLMCQ_F32) {
    e = cudaFuncSetAttribute(radix_hist_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e);
    const int64_t tiles = (numel / 4 + kThreads * 4 - 1) / (kThreads * 4);
    radix_hist_kernel<float><<<stream_grid(tiles, 3), kThreads, smem, st>>>(static_cast<const float*>(x), numel,
                                                                            abs_input, pass, rs, hist);
  } else if (dtype == DLMCQ_BF16) {
    e = cudaFuncSetAttribute(radix_hist_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e);
    const int64_t tiles = (numel / 8 + kThreads * 4 - 1) / (kThreads * 4);
    radix_hist_kernel<__nv_bfloat16><<<stream_grid(tiles, 3), kThreads, smem, st>>>(
        static_cast<const __nv_bfloat16*>(x), numel, abs_input, pass, rs, hist);
  } else {
    return DLMCQ_EINVAL;
  }
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

extern "C" int dlmcq_obs_kth_select(int pass, void* state, void* stream) {
  if (!state || pass < 0 || pass > 2) return DLMCQ_EINVAL;
  const unsigned int* hist = reinterpret_cast<const unsigned int*>(static_cast<char*>(state) + 256);
  radix_select_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(hist, pass, static_cast<RadixState*>(state));
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

extern "C" int dlmcq_obs_kth_values(const void* state, float* values, void* stream) {
  if (!state || !values) return DLMCQ_EINVAL;
  radix_values_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const RadixState*>(state), values);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

extern "C" size_t dlmcq_obs_kth_fast_workspace_bytes(int64_t numel) {
  if (numel < 1) return 0;
  int64_t cap = numel / 16;
  if (cap < 65536) cap = 65536;
  return 512 + static_cast<size_t>(kKthSample) * 4 + 6 * static_cast<size_t>(kRadixBins) * 4 +
         2 * static_cast<size_t>(cap) * 4;
}

extern "C" int dlmcq_obs_kth_fast(const void* x, int64_t numel, int dtype, int flags, int64_t rank0, int64_t rank1,
                                  float* values, int32_t* status, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  if (!x || !values || !status || !workspace || numel < 1 || rank0 < 1 || rank1 < 0) return DLMCQ_EINVAL;
  if (rank0 > numel || rank1 > numel) return DLMCQ_EINVAL;
  if (dtype != DLMCQ_F32 && dtype != DLMCQ_BF16) return DLMCQ_EINVAL;
  if (workspace_bytes < dlmcq_obs_kth_fast_workspace_bytes(numel)) return DLMCQ_EWORKSPACE;
  if (numel < 4 * kKthSample) return DLMCQ_EUNSUPPORTED;          // small tensors: the three-pass select is already cheap
  int64_t cap64 = numel / 16;
  if (cap64 < 65536) cap64 = 65536;
  if (cap64 > 0x7fffffffLL) cap64 = 0x7fffffffLL;
  const unsigned int cap = static_cast<unsigned int>(cap64);
  KthFastState* st = static_cast<KthFastState*>(workspace);
  unsigned int* counter = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + 256);
  unsigned int* sample = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + 512);
  unsigned int* ghist = sample + kKthSample;
  unsigned int* cand = ghist + 6 * kRadixBins;
  cudaStream_t stm = static_cast<cudaStream_t>(stream);
  const int abs_input = (flags & DLMCQ_STATS_ABS_INPUT) ? 1 : 0;
  const unsigned int nr = rank1 > 0 ? 2u : 1u;
  const unsigned long long k0 = static_cast<unsigned long long>(rank0), k1 = static_cast<unsigned long long>(rank1);
  const int64_t vn = dtype == DLMCQ_F32 ? 4 : 8;
  const int grid = stream_grid((numel / vn + kThreads * 4 - 1) / (kThreads * 4), 8);
#define DLMCQ_KTH_LAUNCH(T)                                                                                        \
  do {                                                                                                             \
    const T* xt = static_cast<const T*>(x);                                                                        \
    kth_sample_kernel<T><<<1, kKthCtaThreads, 0, stm>>>(xt, numel, abs_input, k0, k1, nr, st, sample);            \
    DLMCQ_LAUNCH_CHECK();                                                                                          \
    if (abs_input) {                                                                                               \
      if (nr > 1) kth_collect_kernel<T, true, true><<<grid, kThreads, 0, stm>>>(xt, numel, st, cand, cap);         \
      else kth_collect_kernel<T, true, false><<<grid, kThreads, 0, stm>>>(xt, numel, st, cand, cap);               \
    } else {                                                                                                       \
      if (nr > 1) kth_collect_kernel<T, false, true><<<grid, kThreads, 0, stm>>>(xt, numel, st, cand, cap);        \
      else kth_collect_kernel<T, false, false><<<grid, kThreads, 0, stm>>>(xt, numel, st, cand, cap);              \
    }                                                                                                              \
  } while (0)
  if (dtype == DLMCQ_F32) DLMCQ_KTH_LAUNCH(float);
  else DLMCQ_KTH_LAUNCH(__nv_bfloat16);
#undef DLMCQ_KTH_LAUNCH
  DLMCQ_LAUNCH_CHECK();
  // barrier counter + the six global histograms of the cooperative resolve
  cudaError_t e = cudaMemsetAsync(counter, 0, 256, stm);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaMemsetAsync(ghist, 0, 6 * kRadixBins * sizeof(unsigned int), stm);
  if (e != cudaSuccess) return set_cuda_error(e);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_sms());
  cfg.blockDim = dim3(kKthCtaThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stm;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;        // all CTAs co-resident: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define DLMCQ_KTH_RESOLVE(T, ABS)                                                                              \
  cudaLaunchKernelEx(&cfg, kth_resolve_kernel<T, ABS>, static_cast<const KthFastState*>(st),                   \
                     static_cast<const unsigned int*>(cand), cap, ghist, counter, static_cast<const T*>(x), numel, \
                     values, status)
  if (dtype == DLMCQ_F32) e = abs_input ? DLMCQ_KTH_RESOLVE(float, true) : DLMCQ_KTH_RESOLVE(float, false);
  else e = abs_input ? DLMCQ_KTH_RESOLVE(__nv_bfloat16, true) : DLMCQ_KTH_RESOLVE(__nv_bfloat16, false);
#undef DLMCQ_KTH_RESOLVE
  if (e != cudaSuccess) return set_cuda_error(e);
  return DLMCQ_OK;
}

// One call for the whole observer: the one-read path (tensors of >= 65536 elements; exact whether or not its bracket
// held) or, for small tensors, the three-pass select.  No host read anywhere.
extern "C" int dlmcq_obs_kth_auto(const void* x, int64_t numel, int dtype, int flags, int64_t rank0, int64_t rank1,
                                  float* values, int32_t* status, void* state, void* scratch, size_t scratch_bytes,
                                  void* stream) {
  if (!x || !values || !status || !state) return DLMCQ_EINVAL;
  int rc = DLMCQ_EUNSUPPORTED;
  if (scratch && numel >= 4 * kKthSample)
    rc = dlmcq_obs_kth_fast(x, numel, dtype, flags, rank0, rank1, values, status, scratch, scratch_bytes, stream);
  if (rc != DLMCQ_EUNSUPPORTED) return rc;
  cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int32_t), static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_cuda_error(e);
  if ((rc = dlmcq_obs_kth_begin(state, rank0, rank1, stream)) != DLMCQ_OK) return rc;
  for (int pass = 0; pass < 3; ++pass) {
    if ((rc = dlmcq_obs_kth_hist(x, numel, dtype, flags, pass, state, stream)) != DLMCQ_OK) return rc;
    if ((rc = dlmcq_obs_kth_select(pass, state, stream)) != DLMCQ_OK) return rc;
  }
  return dlmcq_obs_kth_values(state, values, stream);
}
