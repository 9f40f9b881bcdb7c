// bnq_kernels.cu - the activation fake-quant fused into its PRODUCER:
//
//     conv output x --BatchNorm--(+ residual)--(ReLU)--> a --QBase fake-quant--> a_q --> next QConv
//
// In the reference this is nn.BatchNorm2d -> (+=) -> nn.ReLU -> QBase.forward's input branch
// (dlmc/quantization/scalar/modules/base.py:96-102, entered from modules/conv.py:13-19): four separate passes over
// the activation per direction (cuDNN batch-norm 12 B/elem, add 12, ReLU 8, the eager fake-quant chain ~70) and
// their autograd mirror images.  Here the whole chain is
//     forward  : one statistics pass (read x: 4 B/elem) + one apply pass (read x [, identity]; write a_q [, a])
//     backward : one reduction pass (d gamma, d beta, d in_scale) + one pass that writes dx
// i.e. 12 + 20 = 32 B/elem for the plain chain including the quantizer, against 52 B/elem for an UN-quantised
// BatchNorm + ReLU in a library implementation.  Nothing here is a contraction: HBM-bound, no tensor cores.
//
// Layout: channels-last.  The tensor is a dense [rows, C] matrix (rows = N*H*W, C innermost), which is what a
// channels_last NCHW tensor or a [N, C] matrix is in memory.  A thread owns one 128-bit vector of channels
// (4 fp32 / 8 bf16) for the whole kernel, so every per-channel quantity (mean, gamma*invstd, beta, the running
// sums) lives in its registers; the txw threads across a row read one contiguous C*sizeof(T) run, `ty` such
// rows are adjacent, and every thread keeps kBnqU independent 128-bit loads in flight (16 KB per CTA).
// Per-channel reductions: per-thread sums -> fixed-order shared-memory fold over the ty rows -> per-CTA partials
// -> a finalisation kernel that adds the partials in double in a fixed order (deterministic, no atomics).
//
// Arithmetic: BatchNorm is floating-point reduction work - results agree with torch's F.batch_norm within
// reduction-order tolerance (stated in tests/test_gpu_bnq.py), not bit for bit.  The quantizer stage is the
// same `fq_vec<FORM_AFFINE>` code as the stand-alone kernels: given the BatchNorm output `a` (which the kernel
// can also write), a_q is BIT-IDENTICAL to dlmcq_fq_forward(a) - that is how parity with the reference's
// quantizer chain is carried over (tests: fused a_q == fq_forward(fused a)).
#include "fq_math.cuh"

namespace dlmcq {

constexpr int kBnqU = 4;          // independent 128-bit loads per thread per pass
constexpr int kBnqThreads = 256;

struct BnqGeom {
  int64_t rows;
  int32_t C;     // channels
  int32_t cv;    // 128-bit vectors per row
  int32_t txw;   // threads across a row (<= 256)
  int32_t ty;    // rows covered by one "thread row" step
  int32_t gy;    // CTAs across a row (cv > 256)
  int32_t nbx;   // CTAs along the rows for the REDUCING kernels (= number of per-channel partials)
};

template <typename T>
static inline BnqGeom make_bnq_geom(int64_t rows, int64_t C) {
  BnqGeom g;
  g.rows = rows;
  g.C = static_cast<int32_t>(C);
  g.cv = static_cast<int32_t>(C / Vec<T>::N);
  g.txw = g.cv < kBnqThreads ? g.cv : kBnqThreads;
  if (g.txw < 1) g.txw = 1;
  g.ty = kBnqThreads / g.txw;
  if (g.ty < 1) g.ty = 1;
  g.gy = (g.cv + g.txw - 1) / g.txw;
  const int64_t rpp = static_cast<int64_t>(g.ty) * kBnqU;
  const int64_t passes = (rows + rpp - 1) / rpp;
  int64_t target = static_cast<int64_t>(num_sms()) * 8 / g.gy;   // ~8 resident CTAs per SM
  if (target < 1) target = 1;
  int64_t cap = rows / 16;   // every CTA leaves 2*C partial floats behind: keep them well below the tensor itself
  if (cap < 1) cap = 1;
  int64_t nbx = passes < target ? passes : target;
  if (nbx > cap) nbx = cap;
  if (nbx < 1) nbx = 1;
  g.nbx = static_cast<int32_t>(nbx);
  return g;
}

// workspace: [header 256 B][partials nbx*2*C][scale partials nbx*gy][coef 2*C]
struct BnqWs {
  float* part;
  float* part_s;
  float* coef;
};
static inline size_t bnq_ws_floats(const BnqGeom& g) {
  return static_cast<size_t>(g.nbx) * 2 * g.C + static_cast<size_t>(g.nbx) * g.gy + 2 * static_cast<size_t>(g.C);
}
static inline BnqWs bnq_ws(void* ws, const BnqGeom& g) {
  BnqWs w;
  w.part = ws_partials(ws);
  w.part_s = w.part + static_cast<size_t>(g.nbx) * 2 * g.C;
  w.coef = w.part_s + static_cast<size_t>(g.nbx) * g.gy;
  return w;
}

// per-thread channel constants: VN consecutive channels
template <int VN>
struct BnCh {
  float mean[VN], a[VN], b[VN], istd[VN];
};
template <int VN>
__device__ __forceinline__ void load_bnch(BnCh<VN>& c, const float* __restrict__ gamma, const float* __restrict__ beta,
                                          const float* __restrict__ mean, const float* __restrict__ invstd, int ch0) {
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    const float gm = gamma ? __ldg(gamma + ch0 + e) : 1.f;
    c.mean[e] = __ldg(mean + ch0 + e);
    c.istd[e] = __ldg(invstd + ch0 + e);
    c.a[e] = gm * c.istd[e];
    c.b[e] = beta ? __ldg(beta + ch0 + e) : 0.f;
  }
}
// z = (x - mean) * (gamma * invstd) + beta  (+ identity);  act = relu(z) or z
template <int VN>
__device__ __forceinline__ void bn_act_vec(const float (&x)[VN], const BnCh<VN>& c, const float* idn, bool relu,
                                           float (&act)[VN]) {
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    float z = __fmaf_rn(x[e] - c.mean[e], c.a[e], c.b[e]);
    if (idn) z = z + idn[e];
    act[e] = relu ? relu_ref(z) : z;
  }
}

#define BNQ_THREAD_COORDS()                                             \
  const int tx = static_cast<int>(threadIdx.x) % gm.txw;               \
  const int ty = static_cast<int>(threadIdx.x) / gm.txw;               \
  const int col = static_cast<int>(blockIdx.y) * gm.txw + tx;          \
  const bool active = (col < gm.cv) && (ty < gm.ty);                   \
  const int64_t rpp = static_cast<int64_t>(gm.ty) * kBnqU

// fold the per-thread sums over the ty thread rows (fixed order) and store this CTA's partials [2][C]
template <int VN>
__device__ __forceinline__ void bnq_store_partials(const float (&s1)[VN], const float (&s2)[VN], float* sm,
                                                   const BnqGeom& gm, int tx, int ty, int col, bool active,
                                                   float* __restrict__ part) {
  float* mine = sm + (static_cast<size_t>(ty) * gm.txw + tx) * (2 * VN);
  if (ty < gm.ty) {
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      mine[e] = s1[e];
      mine[VN + e] = s2[e];
    }
  }
  __syncthreads();
  if (ty == 0 && active) {
    float a1[VN], a2[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) { a1[e] = 0.f; a2[e] = 0.f; }
    for (int r = 0; r < gm.ty; ++r) {
      const float* p = sm + (static_cast<size_t>(r) * gm.txw + tx) * (2 * VN);
#pragma unroll
      for (int e = 0; e < VN; ++e) { a1[e] += p[e]; a2[e] += p[VN + e]; }
    }
    float* o = part + static_cast<size_t>(blockIdx.x) * 2 * gm.C + static_cast<size_t>(col) * VN;
#pragma unroll
    for (int e = 0; e < VN; ++e) { o[e] = a1[e]; o[gm.C + e] = a2[e]; }
  }
}

// ---------------------------------------------------------------------------------------
// forward, pass 1: per-channel sums of (x - k) and (x - k)^2, k = the channel's value in row 0 (a sample of the
// data: removes the cancellation of E[x^2] - mean^2 at no cost)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBnqThreads, 4)
bnq_stats_kernel(const T* __restrict__ x, BnqGeom gm, float* __restrict__ part) {
  using V = Vec<T>;
  using raw = typename V::raw;
  constexpr int VN = V::N;
  extern __shared__ __align__(16) float sm[];
  BNQ_THREAD_COORDS();
  const raw* xv = reinterpret_cast<const raw*>(x);
  float s1[VN], s2[VN], k[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) { s1[e] = 0.f; s2[e] = 0.f; k[e] = 0.f; }
  if (active) {
    const raw r0 = __ldg(xv + col);
    V::unpack(r0, k);
  }
  for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * rpp; r0 < gm.rows; r0 += static_cast<int64_t>(gridDim.x) * rpp) {
    raw r[kBnqU];
    bool ok[kBnqU];
#pragma unroll
    for (int u = 0; u < kBnqU; ++u) {
      const int64_t row = r0 + ty + static_cast<int64_t>(u) * gm.ty;
      ok[u] = active && row < gm.rows;
      if (ok[u]) r[u] = ld_stream(xv + row * gm.cv + col);
    }
#pragma unroll
    for (int u = 0; u < kBnqU; ++u) {
      if (!ok[u]) continue;
      float f[VN];
      V::unpack(r[u], f);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const float d = f[e] - k[e];
        s1[e] += d;
        s2[e] = __fmaf_rn(d, d, s2[e]);
      }
    }
  }
  bnq_store_partials<VN>(s1, s2, sm, gm, tx, ty, col, active, part);
}

// one CTA per 32 channels: 8 thread rows stride over the nbx partials (coalesced over channels), fixed-order fold
template <typename T>
__global__ void __launch_bounds__(256)
bnq_stats_finalize_kernel(const T* __restrict__ x, const float* __restrict__ part, BnqGeom gm, float eps, float momentum,
                          float* __restrict__ running_mean, float* __restrict__ running_var,
                          float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  __shared__ double sh[2][8][32];
  const int cx = threadIdx.x & 31, by = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double a1 = 0.0, a2 = 0.0;
  if (c < gm.C) {
    for (int b = by; b < gm.nbx; b += 8) {
      a1 += static_cast<double>(part[static_cast<size_t>(b) * 2 * gm.C + c]);
      a2 += static_cast<double>(part[static_cast<size_t>(b) * 2 * gm.C + gm.C + c]);
    }
  }
  sh[0][by][cx] = a1;
  sh[1][by][cx] = a2;
  __syncthreads();
  if (by == 0 && c < gm.C) {
    double t1 = 0.0, t2 = 0.0;
    for (int r = 0; r < 8; ++r) { t1 += sh[0][r][cx]; t2 += sh[1][r][cx]; }
    const double n = static_cast<double>(gm.rows);
    const double k = static_cast<double>(to_f32<T>(x[c]));
    const double m1 = t1 / n;
    const double mean = k + m1;
    double var = t2 / n - m1 * m1;              // biased variance of the batch
    if (var < 0.0) var = 0.0;
    save_mean[c] = static_cast<float>(mean);
    save_invstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (running_mean) {
      const double mo = static_cast<double>(momentum);
      running_mean[c] = static_cast<float>((1.0 - mo) * static_cast<double>(running_mean[c]) + mo * mean);
      const double unbiased = n > 1.0 ? var * (n / (n - 1.0)) : var;
      running_var[c] = static_cast<float>((1.0 - mo) * static_cast<double>(running_var[c]) + mo * unbiased);
    }
  }
}

// eval mode: statistics are the running buffers
__global__ void bnq_eval_prepare_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                        int C, float eps, float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    save_mean[c] = running_mean[c];
    save_invstd[c] = 1.0f / sqrtf(running_var[c] + eps);
  }
}

// ---------------------------------------------------------------------------------------
// forward, pass 2: normalise (+ identity) (+ ReLU) -> a (optional) and fake-quant(a) (optional)
// ---------------------------------------------------------------------------------------
template <typename T, bool HAS_ID, bool QUANT>
__global__ void __launch_bounds__(kBnqThreads, 3)
bnq_apply_kernel(const T* __restrict__ x, const T* __restrict__ idn, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                 T* __restrict__ a_out, T* __restrict__ q_out, BnqGeom gm, int relu, const float* __restrict__ scale,
                 const float* __restrict__ offset, float g, float lo, float hi) {
  using V = Vec<T>;
  using raw = typename V::raw;
  constexpr int VN = V::N;
  BNQ_THREAD_COORDS();
  if (!active) return;
  BnCh<VN> c;
  load_bnch<VN>(c, gamma, beta, mean, invstd, col * VN);
  ChanParams p;
  if (QUANT) p = make_params<DLMCQ_FORM_AFFINE>(scale, offset, 0, g, lo, hi);
  const raw* xv = reinterpret_cast<const raw*>(x);
  const raw* iv = reinterpret_cast<const raw*>(idn);
  raw* av = reinterpret_cast<raw*>(a_out);
  raw* qv = reinterpret_cast<raw*>(q_out);
  for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * rpp; r0 < gm.rows; r0 += static_cast<int64_t>(gridDim.x) * rpp) {
    raw rx[kBnqU], ri[kBnqU];
    bool ok[kBnqU];
#pragma unroll
    for (int u = 0; u < kBnqU; ++u) {
      const int64_t row = r0 + ty + static_cast<int64_t>(u) * gm.ty;
      ok[u] = row < gm.rows;
      if (ok[u]) {
        rx[u] = ld_stream(xv + row * gm.cv + col);
        if (HAS_ID) ri[u] = ld_stream(iv + row * gm.cv + col);
      }
    }
#pragma unroll
    for (int u = 0; u < kBnqU; ++u) {
      if (!ok[u]) continue;
      const int64_t idx = (r0 + ty + static_cast<int64_t>(u) * gm.ty) * gm.cv + col;
      float f[VN], fi[VN], act[VN];
      V::unpack(rx[u], f);
      if (HAS_ID) V::unpack(ri[u], fi);
      bn_act_vec<VN>(f, c, HAS_ID ? fi : nullptr, relu != 0, act);
      if (av) {
        // `a` is re-read by the backward pass and by the next block's residual add: default L2 policy
        av[idx] = V::pack(act);
      }
      if (QUANT) {
        float src[VN], code[VN], y[VN];
        if (sizeof(T) == 2) {
          // bf16: the unfused chain quantises the bf16-ROUNDED activation
#pragma unroll
          for (int e = 0; e < VN; ++e) src[e] = to_f32<T>(from_f32<T>(act[e]));
        } else {
#pragma unroll
          for (int e = 0; e < VN; ++e) src[e] = act[e];
        }
        fq_vec<DLMCQ_FORM_AFFINE, VN>(src, p, lo, hi, code, y);
        st_stream(qv + idx, V::pack(y));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward.  dz = d(loss)/d(z), z = BatchNorm output (+ identity) before the ReLU:
//     da = d_a (gradient of the plain output, if any) + fake-quant backward(a, d_q)     [fq_vec_bwd<AFFINE>]
//     dz = da * 1[a > 0] (ReLU) or da
//   RESID = false: `a` is recomputed from x (nothing but x and the upstream gradients is read);
//   RESID = true : z contains the identity, so the saved plain output `a` is read instead, and dz - which IS the
//                  gradient of the identity branch - is written once and re-read by the dx pass.
// pass 1 (reduce): per-channel sum(dz), sum(dz * xhat), and the per-tensor scale-gradient sum.
// ---------------------------------------------------------------------------------------
template <typename T, int VN, bool RESID, bool QUANT>
__device__ __forceinline__ void bnq_dz_vec(const float (&fx)[VN], const float (&fa)[VN], const float* fda,
                                           const float (&fdq)[VN], const BnCh<VN>& c, bool relu, const ChanParams& p,
                                           float lo, float hi, float& acc_s, float (&dz)[VN]) {
  float act[VN];
  if (RESID) {
#pragma unroll
    for (int e = 0; e < VN; ++e) act[e] = fa[e];
  } else {
    bn_act_vec<VN>(fx, c, nullptr, relu, act);
    if (sizeof(T) == 2) {
#pragma unroll
      for (int e = 0; e < VN; ++e) act[e] = to_f32<T>(from_f32<T>(act[e]));
    }
  }
  float da[VN];
  if (QUANT) {
    float dummy = 0.f;
    fq_vec_bwd<DLMCQ_FORM_AFFINE, false, VN>(act, fdq, p, lo, hi, da, acc_s, dummy);
  } else {
#pragma unroll
    for (int e = 0; e < VN; ++e) da[e] = 0.f;
  }
  if (fda) {
#pragma unroll
    for (int e = 0; e < VN; ++e) da[e] += fda[e];
  }
#pragma unroll
  for (int e = 0; e < VN; ++e) dz[e] = relu ? ((act[e] > 0.f) ? da[e] : 0.f) : da[e];
}

template <typename T, bool RESID, bool QUANT>
__global__ void __launch_bounds__(kBnqThreads, 3)
bnq_bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ a_saved, const T* __restrict__ d_a,
                      const T* __restrict__ d_q, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ mean, const float* __restrict__ invstd, T* __restrict__ dz_out,
                      BnqGeom gm, int relu, const float* __restrict__ scale, const float* __restrict__ offset, float g,
                      float lo, float hi, float* __restrict__ part, float* __restrict__ part_s) {
  using V = Vec<T>;
  using raw = typename V::raw;
  constexpr int VN = V::N;
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[64];
  BNQ_THREAD_COORDS();
  BnCh<VN> c;
  if (active) load_bnch<VN>(c, gamma, beta, mean, invstd, col * VN);
  ChanParams p;
  if (QUANT) p = make_params<DLMCQ_FORM_AFFINE>(scale, offset, 0, g, lo, hi);
  const raw* xv = reinterpret_cast<const raw*>(x);
  const raw* av = reinterpret_cast<const raw*>(a_saved);
  const raw* dav = reinterpret_cast<const raw*>(d_a);
  const raw* dqv = reinterpret_cast<const raw*>(d_q);
  raw* zv = reinterpret_cast<raw*>(dz_out);
  float sdb[VN], sdg[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) { sdb[e] = 0.f; sdg[e] = 0.f; }
  float acc[1] = {0.f};
  constexpr int U = RESID ? 2 : kBnqU;      // RESID reads up to four streams per row: two rows in flight are enough
  const int64_t step = static_cast<int64_t>(gm.ty) * U;
  for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * step; r0 < gm.rows; r0 += static_cast<int64_t>(gridDim.x) * step) {
    raw rx[U], ra[U], rda[U], rdq[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = r0 + ty + static_cast<int64_t>(u) * gm.ty;
      ok[u] = active && row < gm.rows;
      if (ok[u]) {
        const int64_t idx = row * gm.cv + col;
        rx[u] = ld_stream(xv + idx);
        if (RESID) ra[u] = ld_stream(av + idx);
        if (d_a) rda[u] = ld_stream(dav + idx);
        if (QUANT) rdq[u] = ld_stream(dqv + idx);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      const int64_t idx = (r0 + ty + static_cast<int64_t>(u) * gm.ty) * gm.cv + col;
      float fx[VN], fa[VN], fda[VN], fdq[VN], dz[VN];
      V::unpack(rx[u], fx);
      if (RESID) V::unpack(ra[u], fa);
      if (d_a) V::unpack(rda[u], fda);
      if (QUANT) V::unpack(rdq[u], fdq);
      bnq_dz_vec<T, VN, RESID, QUANT>(fx, fa, d_a ? fda : nullptr, fdq, c, relu != 0, p, lo, hi, acc[0], dz);
      if (zv) {
        if (sizeof(T) == 2) {   // the dx pass and the identity branch see the rounded value: reduce the same numbers
#pragma unroll
          for (int e = 0; e < VN; ++e) dz[e] = to_f32<T>(from_f32<T>(dz[e]));
        }
        zv[idx] = V::pack(dz);  // re-read by the dx pass (and by the identity branch): default cache policy
      }
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const float xh = (fx[e] - c.mean[e]) * c.istd[e];
        sdb[e] += dz[e];
        sdg[e] = __fmaf_rn(dz[e], xh, sdg[e]);
      }
    }
  }
  bnq_store_partials<VN>(sdb, sdg, sm, gm, tx, ty, col, active, part);
  if (QUANT) {
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) part_s[static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x] = acc[0];
  }
}

// per-channel d gamma / d beta and the coefficients of the dx pass; the last CTA reduces the scale gradient
__global__ void __launch_bounds__(256)
bnq_bwd_finalize_kernel(const float* __restrict__ part, const float* __restrict__ part_s, BnqGeom gm, int n_part_s,
                        int training, float g, float* __restrict__ dgamma, float* __restrict__ dbeta,
                        float* __restrict__ coef, float* __restrict__ dscale) {
  __shared__ double sh[2][8][32];
  const int cx = threadIdx.x & 31, by = threadIdx.x >> 5;
  if (blockIdx.x == gridDim.x - 1) {          // scale gradient: fixed-order sum of the per-CTA partials in double
    double s = 0.0;
    for (int i = threadIdx.x; i < n_part_s; i += blockDim.x) s += static_cast<double>(part_s[i]);
    s = warp_sum(s);
    if (cx == 0) sh[0][by][0] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int r = 0; r < 8; ++r) t += sh[0][r][0];
      if (dscale) dscale[0] = static_cast<float>(t) * g;     // chain through grad_scale (utils.py:24-27)
    }
    return;
  }
  const int c = blockIdx.x * 32 + cx;
  double a1 = 0.0, a2 = 0.0;
  if (c < gm.C) {
    for (int b = by; b < gm.nbx; b += 8) {
      a1 += static_cast<double>(part[static_cast<size_t>(b) * 2 * gm.C + c]);
      a2 += static_cast<double>(part[static_cast<size_t>(b) * 2 * gm.C + gm.C + c]);
    }
  }
  sh[0][by][cx] = a1;
  sh[1][by][cx] = a2;
  __syncthreads();
  if (by == 0 && c < gm.C) {
    double db = 0.0, dg = 0.0;
    for (int r = 0; r < 8; ++r) { db += sh[0][r][cx]; dg += sh[1][r][cx]; }
    if (dbeta) dbeta[c] = static_cast<float>(db);
    if (dgamma) dgamma[c] = static_cast<float>(dg);
    const double n = static_cast<double>(gm.rows);
    coef[c] = training ? static_cast<float>(db / n) : 0.f;            // mean(dz)
    coef[gm.C + c] = training ? static_cast<float>(dg / n) : 0.f;     // mean(dz * xhat)
  }
}

// pass 2: dx = gamma*invstd * (dz - mean(dz) - xhat * mean(dz*xhat))     (eval mode: gamma*invstd * dz)
template <typename T, bool RESID, bool QUANT>
__global__ void __launch_bounds__(kBnqThreads, 3)
bnq_bwd_dx_kernel(const T* __restrict__ x, const T* __restrict__ dz_in, const T* __restrict__ d_a,
                  const T* __restrict__ d_q, const float* __restrict__ gamma, const float* __restrict__ beta,
                  const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ coef,
                  T* __restrict__ dx, BnqGeom gm, int relu, const float* __restrict__ scale,
                  const float* __restrict__ offset, float g, float lo, float hi) {
  using V = Vec<T>;
  using raw = typename V::raw;
  constexpr int VN = V::N;
  BNQ_THREAD_COORDS();
  if (!active) return;
  BnCh<VN> c;
  load_bnch<VN>(c, gamma, beta, mean, invstd, col * VN);
  float c1[VN], c2[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    c1[e] = __ldg(coef + col * VN + e);
    c2[e] = __ldg(coef + gm.C + col * VN + e);
  }
  ChanParams p;
  if (QUANT) p = make_params<DLMCQ_FORM_AFFINE>(scale, offset, 0, g, lo, hi);
  const raw* xv = reinterpret_cast<const raw*>(x);
  const raw* zv = reinterpret_cast<const raw*>(dz_in);
  const raw* dav = reinterpret_cast<const raw*>(d_a);
  const raw* dqv = reinterpret_cast<const raw*>(d_q);
  raw* ov = reinterpret_cast<raw*>(dx);
  for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * rpp; r0 < gm.rows; r0 += static_cast<int64_t>(gridDim.x) * rpp) {
    raw rx[kBnqU], r1[kBnqU], r2[kBnqU];
    bool ok[kBnqU];
#pragma unroll
    for (int u = 0; u < kBnqU; ++u) {
      const int64_t row = r0 + ty + static_cast<int64_t>(u) * gm.ty;
      ok[u] = row < gm.rows;
      if (ok[u]) {
        const int64_t idx = row * gm.cv + col;
        rx[u] = ld_stream(xv + idx);
        if (RESID) {
          r1[u] = ld_stream(zv + idx);
        } else {
          if (d_a) r1[u] = ld_stream(dav + idx);
          if (QUANT) r2[u] = ld_stream(dqv + idx);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kBnqU; ++u) {
      if (!ok[u]) continue;
      const int64_t idx = (r0 + ty + static_cast<int64_t>(u) * gm.ty) * gm.cv + col;
      float fx[VN], dz[VN], o[VN];
      V::unpack(rx[u], fx);
      if (RESID) {
        V::unpack(r1[u], dz);
      } else {
        float fda[VN], fdq[VN], dummy = 0.f;
        if (d_a) V::unpack(r1[u], fda);
        if (QUANT) V::unpack(r2[u], fdq);
        bnq_dz_vec<T, VN, false, QUANT>(fx, fx, d_a ? fda : nullptr, fdq, c, relu != 0, p, lo, hi, dummy, dz);
      }
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const float xh = (fx[e] - c.mean[e]) * c.istd[e];
        const float t = __fmaf_rn(-xh, c2[e], dz[e] - c1[e]);
        o[e] = c.a[e] * t;
      }
      st_stream(ov + idx, V::pack(o));
    }
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static inline int bnq_check(const dlmcq_bnq_desc* d) {
  if (!d || d->rows < 0 || d->channels < 1) return DLMCQ_EINVAL;
  if (d->dtype != DLMCQ_F32 && d->dtype != DLMCQ_BF16) return DLMCQ_EINVAL;
  const int64_t vn = d->dtype == DLMCQ_F32 ? 4 : 8;
  if (d->channels % vn != 0 || d->channels > (int64_t(1) << 20)) return DLMCQ_EUNSUPPORTED;
  return DLMCQ_OK;
}
static inline int apply_grid(const BnqGeom& g) {
  const int64_t rpp = static_cast<int64_t>(g.ty) * kBnqU;
  int64_t passes = (g.rows + rpp - 1) / rpp;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 1024 / g.gy;      // one 16 KB tile per CTA (as fq_fwd_flat)
  if (passes > cap) passes = cap;
  return passes < 1 ? 1 : static_cast<int>(passes);
}
template <int VN>
static inline size_t bnq_smem(const BnqGeom& g) {
  return static_cast<size_t>(g.ty) * g.txw * 2 * VN * sizeof(float);
}

template <typename T>
static int bnq_forward_t(const void* x, const void* idn, const float* gamma, const float* beta, float* rmean,
                         float* rvar, float* smean, float* sinv, void* a_out, void* q_out, const dlmcq_bnq_desc* d,
                         const dlmcq_qparams* qp, void* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int VN = Vec<T>::N;
  const BnqGeom g = make_bnq_geom<T>(d->rows, d->channels);
  if (ws_bytes < kWsHeaderBytes + bnq_ws_floats(g) * sizeof(float)) return DLMCQ_EWORKSPACE;
  const BnqWs w = bnq_ws(ws, g);
  const bool training = (d->flags & DLMCQ_BNQ_TRAINING) != 0;
  const T* xt = static_cast<const T*>(x);
  if (training) {
    bnq_stats_kernel<T><<<dim3(g.nbx, g.gy), kBnqThreads, bnq_smem<VN>(g), st>>>(xt, g, w.part);
    DLMCQ_LAUNCH_CHECK();
    bnq_stats_finalize_kernel<T><<<(g.C + 31) / 32, 256, 0, st>>>(xt, w.part, g, d->eps, d->momentum, rmean, rvar,
                                                                    smean, sinv);
  } else {
    bnq_eval_prepare_kernel<<<(g.C + 255) / 256, 256, 0, st>>>(rmean, rvar, g.C, d->eps, smean, sinv);
  }
  DLMCQ_LAUNCH_CHECK();
  if (!a_out && !q_out) return DLMCQ_OK;     // statistics only
  const bool quant = q_out != nullptr;
  const int relu = (d->flags & DLMCQ_BNQ_RELU) ? 1 : 0;
  const float lo = quant ? static_cast<float>(qp->lo) : 0.f, hi = quant ? static_cast<float>(qp->hi) : 0.f;
  const float* sc = quant ? qp->scale : nullptr;
  const float* of = quant ? qp->offset : nullptr;
  const float gq = quant ? qp->g : 0.f;
  const dim3 grid(apply_grid(g), g.gy);
  auto k = idn ? (quant ? bnq_apply_kernel<T, true, true> : bnq_apply_kernel<T, true, false>)
               : (quant ? bnq_apply_kernel<T, false, true> : bnq_apply_kernel<T, false, false>);
  k<<<grid, kBnqThreads, 0, st>>>(xt, static_cast<const T*>(idn), gamma, beta, smean, sinv, static_cast<T*>(a_out),
                                  static_cast<T*>(q_out), g, relu, sc, of, gq, lo, hi);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

template <typename T>
static int bnq_backward_t(const void* x, const void* a_saved, const void* d_a, const void* d_q, const float* gamma,
                          const float* beta, const float* smean, const float* sinv, void* dx, void* dz_out,
                          float* dgamma, float* dbeta, float* dscale, const dlmcq_bnq_desc* d, const dlmcq_qparams* qp,
                          void* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int VN = Vec<T>::N;
  const BnqGeom g = make_bnq_geom<T>(d->rows, d->channels);
  if (ws_bytes < kWsHeaderBytes + bnq_ws_floats(g) * sizeof(float)) return DLMCQ_EWORKSPACE;
  const BnqWs w = bnq_ws(ws, g);
  const bool resid = (d->flags & DLMCQ_BNQ_RESIDUAL) != 0;
  const bool quant = d_q != nullptr;
  const int relu = (d->flags & DLMCQ_BNQ_RELU) ? 1 : 0;
  const int training = (d->flags & DLMCQ_BNQ_TRAINING) ? 1 : 0;
  const float lo = quant ? static_cast<float>(qp->lo) : 0.f, hi = quant ? static_cast<float>(qp->hi) : 0.f;
  const float* sc = quant ? qp->scale : nullptr;
  const float* of = quant ? qp->offset : nullptr;
  const float gq = quant ? qp->g : 0.f;
  const T* xt = static_cast<const T*>(x);
  const T* at = static_cast<const T*>(a_saved);
  const T* dat = static_cast<const T*>(d_a);
  const T* dqt = static_cast<const T*>(d_q);
  {
    auto k = resid ? (quant ? bnq_bwd_reduce_kernel<T, true, true> : bnq_bwd_reduce_kernel<T, true, false>)
                   : (quant ? bnq_bwd_reduce_kernel<T, false, true> : bnq_bwd_reduce_kernel<T, false, false>);
    k<<<dim3(g.nbx, g.gy), kBnqThreads, bnq_smem<VN>(g), st>>>(xt, at, dat, dqt, gamma, beta, smean, sinv,
                                                               static_cast<T*>(dz_out), g, relu, sc, of, gq, lo, hi,
                                                               w.part, w.part_s);
    DLMCQ_LAUNCH_CHECK();
  }
  bnq_bwd_finalize_kernel<<<(g.C + 31) / 32 + 1, 256, 0, st>>>(w.part, w.part_s, g, quant ? g.nbx * g.gy : 0, training,
                                                                 gq, dgamma, dbeta, w.coef, quant ? dscale : nullptr);
  DLMCQ_LAUNCH_CHECK();
  if (dx) {
    const dim3 grid(apply_grid(g), g.gy);
    auto k = resid ? bnq_bwd_dx_kernel<T, true, false>
                   : (quant ? bnq_bwd_dx_kernel<T, false, true> : bnq_bwd_dx_kernel<T, false, false>);
    k<<<grid, kBnqThreads, 0, st>>>(xt, static_cast<const T*>(dz_out), dat, dqt, gamma, beta, smean, sinv, w.coef,
                                    static_cast<T*>(dx), g, relu, sc, of, gq, lo, hi);
    DLMCQ_LAUNCH_CHECK();
  }
  return DLMCQ_OK;
}

}  // namespace dlmcq

using namespace dlmcq;

extern "C" size_t dlmcq_bnq_workspace_bytes(const dlmcq_bnq_desc* d) {
  if (bnq_check(d) != DLMCQ_OK) return 0;
  const BnqGeom g = d->dtype == DLMCQ_F32 ? make_bnq_geom<float>(d->rows, d->channels)
                                          : make_bnq_geom<__nv_bfloat16>(d->rows, d->channels);
  return kWsHeaderBytes + bnq_ws_floats(g) * sizeof(float);
}

extern "C" int dlmcq_bnq_forward(const void* x, const void* identity, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                                 void* a_out, void* q_out, const dlmcq_bnq_desc* desc, const dlmcq_qparams* qp,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = bnq_check(desc)) return e;
  if (desc->rows == 0) return DLMCQ_OK;
  if (!x || !save_mean || !save_invstd || !workspace) return DLMCQ_EINVAL;
  const bool training = (desc->flags & DLMCQ_BNQ_TRAINING) != 0;
  if (!training && (!running_mean || !running_var)) return DLMCQ_EINVAL;
  if ((running_mean == nullptr) != (running_var == nullptr)) return DLMCQ_EINVAL;
  if (q_out && (!qp || !qp->scale || qp->form != DLMCQ_FORM_AFFINE)) return DLMCQ_EINVAL;
  if (((desc->flags & DLMCQ_BNQ_RESIDUAL) != 0) != (identity != nullptr)) return DLMCQ_EINVAL;
  if (!aligned16(x) || !aligned16(identity) || !aligned16(a_out) || !aligned16(q_out)) return DLMCQ_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return desc->dtype == DLMCQ_F32
             ? bnq_forward_t<float>(x, identity, gamma, beta, running_mean, running_var, save_mean, save_invstd, a_out,
                                    q_out, desc, qp, workspace, workspace_bytes, st)
             : bnq_forward_t<__nv_bfloat16>(x, identity, gamma, beta, running_mean, running_var, save_mean, save_invstd,
                                            a_out, q_out, desc, qp, workspace, workspace_bytes, st);
}

extern "C" int dlmcq_bnq_backward(const void* x, const void* a_saved, const void* d_a, const void* d_q,
                                  const float* gamma, const float* beta, const float* save_mean,
                                  const float* save_invstd, void* dx, void* dz_out, float* dgamma, float* dbeta,
                                  float* dscale, const dlmcq_bnq_desc* desc, const dlmcq_qparams* qp, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (int e = bnq_check(desc)) return e;
  if (desc->rows == 0) return DLMCQ_OK;
  if (!x || !save_mean || !save_invstd || !workspace || (!d_a && !d_q)) return DLMCQ_EINVAL;
  const bool resid = (desc->flags & DLMCQ_BNQ_RESIDUAL) != 0;
  if (resid && (!a_saved || !dz_out)) return DLMCQ_EINVAL;
  if (d_q && (!qp || !qp->scale || !dscale || qp->form != DLMCQ_FORM_AFFINE)) return DLMCQ_EINVAL;
  if (!aligned16(x) || !aligned16(a_saved) || !aligned16(d_a) || !aligned16(d_q) || !aligned16(dx) || !aligned16(dz_out))
    return DLMCQ_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return desc->dtype == DLMCQ_F32
             ? bnq_backward_t<float>(x, a_saved, d_a, d_q, gamma, beta, save_mean, save_invstd, dx, dz_out, dgamma, dbeta,
                                     dscale, desc, qp, workspace, workspace_bytes, st)
             : bnq_backward_t<__nv_bfloat16>(x, a_saved, d_a, d_q, gamma, beta, save_mean, save_invstd, dx, dz_out,
                                             dgamma, dbeta, dscale, desc, qp, workspace, workspace_bytes, st);
}
