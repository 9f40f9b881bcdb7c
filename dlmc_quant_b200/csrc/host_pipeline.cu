// host_pipeline.cu - end-to-end entries for HOST-resident tensors, on a caller-owned context.
//
// A caller that owns (pinned) host buffers - the situation of a plugin bound over ctypes/cgo/JNI with no device
// tensors of its own - gets fake-quant forward + backward in one call: the tensor is cut into chunks and each chunk
// flows H2D -> kernel(s) -> D2H on one of kSlots streams, so copies in both directions overlap the kernels.
//
// All state (streams, events, device staging, cursors) lives in a `dlmcq_host_ctx` that the caller creates, passes
// to every call and destroys: there is no process-global state, two contexts never share anything, and a context is
// safe to use from one thread at a time (calls on one context are serialised by its own mutex).
//
// Two result formats:
//   * full     (dlmcq_host_ctx_fq_forward_backward): y and dx come back in the tensor's dtype - 16 B/element over PCIe
//              for fp32 (8 in, 8 out);
//   * compact  (dlmcq_host_ctx_fq_codes): the forward result comes back as the integer CODES (one byte, or two 4-bit
//              codes per byte) and the backward result as a KEEP bit per element.  Both are lossless: y = code*s'+off
//              (dlmcq_import_codes reproduces dlmcq_fq_forward's y bit for bit) and dx = keep ? dy : 0 with the dy the
//              caller already holds.  8.625 B/element over PCIe instead of 16 - the path is PCIe-bound, so that is
//              the speed-up - and one fused kernel per chunk (x and dy read once) instead of two.
// Calls return when the work is enqueued; successive tensors (the layers of a model) keep the copy engines busy and
// dlmcq_host_ctx_synchronize() waits once.  Per-call scale gradients are reduced on the device in a fixed order.
#include <mutex>
#include <new>

#include "fq_math.cuh"

struct dlmcq_host_ctx {
  static constexpr int kSlots = 4;
  int device = 0;
  int64_t chunk = 0;
  cudaStream_t s[kSlots] = {};
  cudaStream_t epi = nullptr;
  cudaEvent_t slot_done[kSlots] = {};
  char* staging = nullptr;
  size_t slot_bytes = 0, buf = 0, wsb = 0;
  float* d_chunk = nullptr;   // per-chunk scale-gradient ring
  float* d_call = nullptr;    // per-call result ring
  int64_t chunk_cursor = 0, call_cursor = 0, rr = 0, in_flight_chunks = 0;
  std::mutex m;
};

namespace dlmcq {

constexpr size_t kAlign = 256;
constexpr int64_t kChunkRing = 65536;   // per-chunk scale-gradient slots (ring)
constexpr int64_t kCallRing = 4096;     // per-call result slots (ring)

static inline size_t up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

// fixed-order sum (in double) of the per-chunk scale gradients of one call: ring indices start .. start+n-1
__global__ void __launch_bounds__(32)
sum_chunks_kernel(const float* __restrict__ ring, int64_t start, int64_t n, int64_t ring_size, float* __restrict__ out) {
  double t = 0.0;
  for (int64_t c = threadIdx.x; c < n; c += 32) t += static_cast<double>(ring[(start + c) % ring_size]);
  t = warp_sum(t);
  if (threadIdx.x == 0) out[0] = static_cast<float>(t);
}

__global__ void set_params_kernel(float* __restrict__ p, float scale, float offset) {
  p[0] = scale;
  p[1] = offset;
}

// Compact chunk kernel: one read of (x, dy) -> packed codes, keep bits, per-CTA scale-gradient partial.
// Every thread owns 8 consecutive elements: 8 code bytes (or 4 when PACK4) and one keep byte.
template <int FORM, typename T, bool PACK4, bool BWD>
__global__ void __launch_bounds__(kThreads, 4)
host_codes_kernel(const T* __restrict__ x, const T* __restrict__ dy, uint8_t* __restrict__ codes,
                  uint8_t* __restrict__ keep, int64_t n, const float* __restrict__ params /* scale, offset */, float g,
                  float lo, float hi, float* __restrict__ dscale, void* ws) {
  __shared__ __align__(16) float smem[64];
  const ChanParams p = make_params<FORM>(params, params + 1, 0, g, lo, hi);
  float acc[1] = {0.f};
  float acc_o = 0.f;
  const int64_t groups = (n + 7) / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; gi < groups; gi += stride) {
    const int64_t i0 = gi * 8;
    float fx[8], fg[8], code[8], y[8], dxv[8];
    if (i0 + 8 <= n) {
      if (sizeof(T) == 4) {
        float a[4], b[4];
        Vec<float>::unpack(ld_stream(reinterpret_cast<const float4*>(x) + 2 * gi), a);
        Vec<float>::unpack(ld_stream(reinterpret_cast<const float4*>(x) + 2 * gi + 1), b);
#pragma unroll
        for (int e = 0; e < 4; ++e) { fx[e] = a[e]; fx[4 + e] = b[e]; }
        if (BWD) {
          Vec<float>::unpack(ld_stream(reinterpret_cast<const float4*>(dy) + 2 * gi), a);
          Vec<float>::unpack(ld_stream(reinterpret_cast<const float4*>(dy) + 2 * gi + 1), b);
#pragma unroll
          for (int e = 0; e < 4; ++e) { fg[e] = a[e]; fg[4 + e] = b[e]; }
        }
      } else {
        Vec<__nv_bfloat16>::unpack(ld_stream(reinterpret_cast<const uint4*>(x) + gi), fx);
        if (BWD) Vec<__nv_bfloat16>::unpack(ld_stream(reinterpret_cast<const uint4*>(dy) + gi), fg);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        fx[e] = i0 + e < n ? to_f32<T>(x[i0 + e]) : 0.f;
        fg[e] = (BWD && i0 + e < n) ? to_f32<T>(dy[i0 + e]) : 0.f;
      }
    }
    fq_vec<FORM, 8>(fx, p, lo, hi, code, y);
    uint32_t kbits = 0;
    if (BWD) {
      fq_vec_bwd<FORM, false, 8>(fx, fg, p, lo, hi, dxv, acc[0], acc_o);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        // dx is dy or 0: "dx != 0 or NaN" reproduces it from the dy the caller holds (a zero dy gives zero either way)
        kbits |= ((dxv[e] != 0.f || dxv[e] != dxv[e]) && i0 + e < n ? 1u : 0u) << e;
      }
    }
    uint32_t b[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = (code[e] == code[e]) ? static_cast<int>(code[e]) : 0;      // NaN code -> 0
      b[e] = static_cast<uint32_t>(c) & (PACK4 ? 0xFu : 0xFFu);
    }
    if (PACK4) {
      uint32_t w = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) w |= b[e] << (4 * e);
      if (i0 + 8 <= n) {
        reinterpret_cast<uint32_t*>(codes)[gi] = w;
      } else {
        const int64_t nb = (n + 1) / 2;
        for (int k = 0; k < 4; ++k)
          if (gi * 4 + k < nb) codes[gi * 4 + k] = static_cast<uint8_t>(w >> (8 * k));
      }
    } else {
      if (i0 + 8 <= n) {
        const uint2 w = make_uint2(b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24),
                                   b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24));
        reinterpret_cast<uint2*>(codes)[gi] = w;
      } else {
        for (int e = 0; e < 8; ++e)
          if (i0 + e < n) codes[i0 + e] = static_cast<uint8_t>(b[e]);
      }
    }
    if (BWD) keep[gi] = static_cast<uint8_t>(kbits);
  }
  if (BWD) {
    block_sum<1>(acc, smem);
    float* partials = ws_partials(ws);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc[0];
    if (take_last_ticket(ws_counter(ws), gridDim.x)) {
      double s = 0.0;
      for (int b2 = threadIdx.x; b2 < static_cast<int>(gridDim.x); b2 += blockDim.x) s += static_cast<double>(partials[b2]);
      s = warp_sum(s);
      double* sm = reinterpret_cast<double*>(smem);
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      __syncthreads();
      if (lane == 0) sm[warp] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += sm[w];
        float r = static_cast<float>(t);
        if (FORM == DLMCQ_FORM_AFFINE || FORM == DLMCQ_FORM_A1) r = r * g;
        dscale[0] = r;
        *ws_counter(ws) = 0u;
      }
    }
  }
}

template <typename T, bool PACK4, bool BWD>
static int launch_codes(int form, const void* x, const void* dy, void* codes, void* keep, int64_t n, const float* params,
                        float g, int lo, int hi, float* dscale, void* ws, cudaStream_t st) {
  const int64_t groups = (n + 7) / 8;
  int64_t grid = (groups + kThreads - 1) / kThreads;
  if (grid > 2048) grid = 2048;
  if (grid < 1) grid = 1;
  const float flo = static_cast<float>(lo), fhi = static_cast<float>(hi);
  const T* xt = static_cast<const T*>(x);
  const T* gt = static_cast<const T*>(dy);
  uint8_t* ct = static_cast<uint8_t*>(codes);
  uint8_t* kt = static_cast<uint8_t*>(keep);
  switch (form) {
    case DLMCQ_FORM_A1:
      host_codes_kernel<DLMCQ_FORM_A1, T, PACK4, BWD><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(xt, gt, ct, kt, n, params, g, flo, fhi, dscale, ws);
      break;
    case DLMCQ_FORM_AFFINE:
      host_codes_kernel<DLMCQ_FORM_AFFINE, T, PACK4, BWD><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(xt, gt, ct, kt, n, params, g, flo, fhi, dscale, ws);
      break;
    case DLMCQ_FORM_ZP:
      host_codes_kernel<DLMCQ_FORM_ZP, T, PACK4, BWD><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(xt, gt, ct, kt, n, params, g, flo, fhi, dscale, ws);
      break;
    case DLMCQ_FORM_SYM:
      host_codes_kernel<DLMCQ_FORM_SYM, T, PACK4, BWD><<<static_cast<unsigned>(grid), kThreads, 0, st>>>(xt, gt, ct, kt, n, params, g, flo, fhi, dscale, ws);
      break;
    default:
      return DLMCQ_EINVAL;
  }
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

static int sync_ctx(dlmcq_host_ctx* c) {
  int status = DLMCQ_OK;
  for (int s = 0; s < dlmcq_host_ctx::kSlots; ++s) {
    cudaError_t e = cudaStreamSynchronize(c->s[s]);
    if (e != cudaSuccess && status == DLMCQ_OK) status = set_cuda_error(e);
  }
  cudaError_t e = cudaStreamSynchronize(c->epi);
  if (e != cudaSuccess && status == DLMCQ_OK) status = set_cuda_error(e);
  c->in_flight_chunks = 0;
  return status;
}

// RAII: make the context's device current for the duration of a call
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

enum class HostMode { kFull, kCodes };

static int host_call(dlmcq_host_ctx* c, HostMode mode, const void* x_host, const void* dy_host, void* out0_host,
                     void* out1_host, float* dscale_host, int64_t numel, int dtype, int form, int lo, int hi, float g,
                     float scale, float offset, int pack4) {
  if (!c || !x_host || !out0_host || numel < 0) return DLMCQ_EINVAL;
  if (dtype != DLMCQ_F32 && dtype != DLMCQ_BF16) return DLMCQ_EINVAL;
  const bool bwd = dy_host != nullptr;
  if (mode == HostMode::kFull && (!bwd || !out1_host || !dscale_host)) return DLMCQ_EINVAL;
  if (bwd && (!out1_host || !dscale_host)) return DLMCQ_EINVAL;
  if (pack4 && (lo < -8 || hi > 15 || (lo < 0 && hi > 7))) return DLMCQ_EINVAL;      // does not fit 4 bits
  if (!pack4 && (lo < -128 || hi > 255 || (lo < 0 && hi > 127))) return DLMCQ_EINVAL;
  const int64_t chunk = c->chunk;
  const int64_t nchunks = (numel + chunk - 1) / chunk;
  if (nchunks > kChunkRing / 2) return DLMCQ_EUNSUPPORTED;
  std::lock_guard<std::mutex> lk(c->m);
  DeviceGuard dg(c->device);
  // ring hygiene: never let un-synchronised calls wrap onto entries that may still be read
  if (c->in_flight_chunks + nchunks > kChunkRing / 2 || c->call_cursor % kCallRing == kCallRing - 1) {
    if (int st = sync_ctx(c)) return st;
  }
  const size_t es = dtype == DLMCQ_F32 ? 4 : 2;
  const size_t buf = c->buf;
  cudaError_t e;
  const int64_t start = c->chunk_cursor;
  bool used[dlmcq_host_ctx::kSlots] = {false};
  int status = DLMCQ_OK;
  for (int64_t ci = 0; ci < nchunks && status == DLMCQ_OK; ++ci) {
    const int s = static_cast<int>(c->rr++ % dlmcq_host_ctx::kSlots);
    cudaStream_t st = c->s[s];
    char* sb = c->staging + s * c->slot_bytes;
    void *dxin = sb, *ddy = sb + buf, *o0 = sb + 2 * buf, *o1 = sb + 3 * buf, *ws = sb + 4 * buf;
    float* d_params = reinterpret_cast<float*>(sb + 4 * buf + c->wsb);
    if (!used[s]) {
      // this call's qparams travel as KERNEL ARGUMENTS of a one-thread fill kernel: they are captured at launch, so
      // no host memory is referenced after this function returns; ordered after the slot's earlier kernels by the stream
      set_params_kernel<<<1, 1, 0, st>>>(d_params, scale, offset);
      if ((e = cudaGetLastError()) != cudaSuccess) { status = set_cuda_error(e); break; }
      used[s] = true;
    }
    const int64_t off = ci * chunk;
    const int64_t len = (numel - off) < chunk ? (numel - off) : chunk;
    const size_t bytes = static_cast<size_t>(len) * es;
    const char* xh = static_cast<const char*>(x_host) + off * es;
    if ((e = cudaMemcpyAsync(dxin, xh, bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
    if (bwd) {
      const char* gh = static_cast<const char*>(dy_host) + off * es;
      if ((e = cudaMemcpyAsync(ddy, gh, bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
    }
    float* ds_slot = c->d_chunk + (start + ci) % kChunkRing;
    if (mode == HostMode::kFull) {
      dlmcq_qparams qp;
      qp.form = form; qp.lo = lo; qp.hi = hi; qp.g = g; qp.scale = d_params; qp.offset = d_params + 1;
      dlmcq_layout l = {1, 1, len, dtype};
      status = dlmcq_fq_forward(dxin, o0, nullptr, &l, &qp, st);
      if (status != DLMCQ_OK) break;
      if ((e = cudaMemcpyAsync(static_cast<char*>(out0_host) + off * es, o0, bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
      status = dlmcq_fq_backward(dxin, ddy, o1, ds_slot, nullptr, &l, &qp, ws, c->wsb, st);
      if (status != DLMCQ_OK) break;
      if ((e = cudaMemcpyAsync(static_cast<char*>(out1_host) + off * es, o1, bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
    } else {
      // chunk boundaries are multiples of 8 elements: whole code words and keep bytes per chunk
      if (dtype == DLMCQ_F32) {
        status = pack4 ? (bwd ? launch_codes<float, true, true>(form, dxin, ddy, o0, o1, len, d_params, g, lo, hi, ds_slot, ws, st)
                              : launch_codes<float, true, false>(form, dxin, nullptr, o0, nullptr, len, d_params, g, lo, hi, nullptr, ws, st))
                       : (bwd ? launch_codes<float, false, true>(form, dxin, ddy, o0, o1, len, d_params, g, lo, hi, ds_slot, ws, st)
                              : launch_codes<float, false, false>(form, dxin, nullptr, o0, nullptr, len, d_params, g, lo, hi, nullptr, ws, st));
      } else {
        status = pack4 ? (bwd ? launch_codes<__nv_bfloat16, true, true>(form, dxin, ddy, o0, o1, len, d_params, g, lo, hi, ds_slot, ws, st)
                              : launch_codes<__nv_bfloat16, true, false>(form, dxin, nullptr, o0, nullptr, len, d_params, g, lo, hi, nullptr, ws, st))
                       : (bwd ? launch_codes<__nv_bfloat16, false, true>(form, dxin, ddy, o0, o1, len, d_params, g, lo, hi, ds_slot, ws, st)
                              : launch_codes<__nv_bfloat16, false, false>(form, dxin, nullptr, o0, nullptr, len, d_params, g, lo, hi, nullptr, ws, st));
      }
      if (status != DLMCQ_OK) break;
      const size_t cbytes = pack4 ? static_cast<size_t>((len + 1) / 2) : static_cast<size_t>(len);
      const size_t coff = pack4 ? static_cast<size_t>(off / 2) : static_cast<size_t>(off);
      if ((e = cudaMemcpyAsync(static_cast<char*>(out0_host) + coff, o0, cbytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
      if (bwd) {
        if ((e = cudaMemcpyAsync(static_cast<char*>(out1_host) + off / 8, o1, static_cast<size_t>((len + 7) / 8), cudaMemcpyDeviceToHost, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
      }
    }
  }
  if (status != DLMCQ_OK) return status;
  if (!bwd) return DLMCQ_OK;
  c->chunk_cursor = (start + nchunks) % kChunkRing;
  c->in_flight_chunks += nchunks;
  // epilogue: wait for this call's chunks on every slot, reduce, ship the scalar home
  for (int s = 0; s < dlmcq_host_ctx::kSlots; ++s) {
    if (!used[s]) continue;
    cudaEventRecord(c->slot_done[s], c->s[s]);
    cudaStreamWaitEvent(c->epi, c->slot_done[s], 0);
  }
  float* out = c->d_call + (c->call_cursor++ % kCallRing);
  sum_chunks_kernel<<<1, 32, 0, c->epi>>>(c->d_chunk, start, nchunks, kChunkRing, out);
  DLMCQ_LAUNCH_CHECK();
  if ((e = cudaMemcpyAsync(dscale_host, out, sizeof(float), cudaMemcpyDeviceToHost, c->epi)) != cudaSuccess)
    return set_cuda_error(e);
  return DLMCQ_OK;
}

}  // namespace dlmcq

using namespace dlmcq;

extern "C" int dlmcq_host_ctx_create(dlmcq_host_ctx** out, int64_t chunk_elems) {
  if (!out || chunk_elems < 8 || chunk_elems % 8 != 0) return DLMCQ_EINVAL;   // multiples of 8: aligned chunks, whole code words
  dlmcq_host_ctx* c = new (std::nothrow) dlmcq_host_ctx();
  if (!c) return DLMCQ_EINVAL;
  cudaError_t e = cudaGetDevice(&c->device);
  if (e != cudaSuccess) { delete c; return set_cuda_error(e); }
  c->chunk = chunk_elems;
  c->buf = up(static_cast<size_t>(chunk_elems) * 4);
  c->wsb = up(dlmcq_workspace_bytes(nullptr));
  c->slot_bytes = 4 * c->buf + c->wsb + up(64);
  const size_t total = dlmcq_host_ctx::kSlots * c->slot_bytes + up(static_cast<size_t>(kChunkRing) * sizeof(float)) +
                       up(static_cast<size_t>(kCallRing) * sizeof(float));
  bool ok = true;
  for (int i = 0; i < dlmcq_host_ctx::kSlots && ok; ++i) {
    ok = cudaStreamCreateWithFlags(&c->s[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->slot_done[i], cudaEventDisableTiming) == cudaSuccess;
  }
  ok = ok && cudaStreamCreateWithFlags(&c->epi, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&c->staging), total) == cudaSuccess;
  ok = ok && cudaMemset(c->staging, 0, total) == cudaSuccess;      // ticket counters of the per-slot workspaces start at 0
  if (!ok) {
    const int st = set_cuda_error(cudaGetLastError());
    dlmcq_host_ctx_destroy(c);
    return st;
  }
  c->d_chunk = reinterpret_cast<float*>(c->staging + dlmcq_host_ctx::kSlots * c->slot_bytes);
  c->d_call = reinterpret_cast<float*>(c->staging + dlmcq_host_ctx::kSlots * c->slot_bytes +
                                       up(static_cast<size_t>(kChunkRing) * sizeof(float)));
  *out = c;
  return DLMCQ_OK;
}

extern "C" int dlmcq_host_ctx_destroy(dlmcq_host_ctx* c) {
  if (!c) return DLMCQ_OK;
  {
    DeviceGuard dg(c->device);
    for (int i = 0; i < dlmcq_host_ctx::kSlots; ++i) {
      if (c->s[i]) { cudaStreamSynchronize(c->s[i]); cudaStreamDestroy(c->s[i]); }
      if (c->slot_done[i]) cudaEventDestroy(c->slot_done[i]);
    }
    if (c->epi) { cudaStreamSynchronize(c->epi); cudaStreamDestroy(c->epi); }
    if (c->staging) cudaFree(c->staging);
  }
  delete c;
  return DLMCQ_OK;
}

extern "C" int dlmcq_host_ctx_synchronize(dlmcq_host_ctx* c) {
  if (!c) return DLMCQ_EINVAL;
  std::lock_guard<std::mutex> lk(c->m);
  DeviceGuard dg(c->device);
  return sync_ctx(c);
}

extern "C" int dlmcq_host_ctx_fq_forward_backward(dlmcq_host_ctx* ctx, const void* x_host, const void* dy_host,
                                                  void* y_host, void* dx_host, float* dscale_host, int64_t numel,
                                                  int dtype, int form, int lo, int hi, float g, float scale,
                                                  float offset) {
  return host_call(ctx, HostMode::kFull, x_host, dy_host, y_host, dx_host, dscale_host, numel, dtype, form, lo, hi, g,
                   scale, offset, 0);
}

extern "C" int dlmcq_host_ctx_fq_codes(dlmcq_host_ctx* ctx, const void* x_host, const void* dy_host, void* codes_host,
                                       void* keep_host, float* dscale_host, int64_t numel, int dtype, int form, int lo,
                                       int hi, float g, float scale, float offset, int pack4) {
  return host_call(ctx, HostMode::kCodes, x_host, dy_host, codes_host, keep_host, dscale_host, numel, dtype, form, lo, hi,
                   g, scale, offset, pack4 ? 1 : 0);
}
