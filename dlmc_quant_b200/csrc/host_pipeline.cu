// host_pipeline.cu - end-to-end entry for HOST-resident tensors.
//
// A caller that owns (pinned) host buffers - the situation of a plugin bound over ctypes/cgo/JNI
// with no device tensors of its own - gets fake-quant forward + backward in one call: the tensor
// is cut into chunks and each chunk flows H2D -> forward kernel -> backward kernel -> D2H on one
// of kSlots streams, so copies in both directions overlap the kernels.  The arithmetic is the
// same dlmcq_fq_forward / dlmcq_fq_backward code path (per-tensor qparams).
//
// The *_async form returns as soon as the work is enqueued: successive tensors (the layers of a
// model) keep the pipeline full instead of draining it at every call; dlmcq_host_synchronize()
// waits once.  Per-call scale gradients are reduced on the device (fixed order) by an epilogue
// stream that waits for the call's last chunk on every slot.
#include <mutex>

#include "common.cuh"

namespace dlmcq {

constexpr int kSlots = 4;
constexpr size_t kAlign = 256;
constexpr int64_t kChunkRing = 65536;   // per-chunk scale-gradient slots (ring)
constexpr int64_t kCallRing = 4096;     // per-call result slots (ring)

static inline size_t up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct StreamPool {
  cudaStream_t s[kSlots];
  cudaStream_t epi;
  cudaEvent_t slot_done[kSlots];
  int64_t chunk_cursor = 0;   // next free entry of the per-chunk ring
  int64_t call_cursor = 0;
  int64_t rr = 0;             // round-robin slot counter, continues across calls
  int64_t in_flight_chunks = 0;
  bool ok = false;
};
static StreamPool g_pools[64];
static std::mutex g_pool_mutex;

static StreamPool* get_pool() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  StreamPool& p = g_pools[dev];
  if (!p.ok) {
    for (int i = 0; i < kSlots; ++i) {
      if (cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&p.slot_done[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaStreamCreateWithFlags(&p.epi, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    p.ok = true;
  }
  return &p;
}

static int sync_pool(StreamPool* p) {
  int status = DLMCQ_OK;
  for (int s = 0; s < kSlots; ++s) {
    cudaError_t e = cudaStreamSynchronize(p->s[s]);
    if (e != cudaSuccess && status == DLMCQ_OK) status = set_cuda_error(e);
  }
  cudaError_t e = cudaStreamSynchronize(p->epi);
  if (e != cudaSuccess && status == DLMCQ_OK) status = set_cuda_error(e);
  p->in_flight_chunks = 0;
  return status;
}

// fixed-order sum of the per-chunk scale gradients of one call (ring indices start .. start+n-1)
__global__ void sum_chunks_kernel(const float* __restrict__ ring, int64_t start, int64_t n, int64_t ring_size,
                                  float* __restrict__ out) {
  double t = 0.0;
  for (int64_t c = 0; c < n; ++c) t += static_cast<double>(ring[(start + c) % ring_size]);
  out[0] = static_cast<float>(t);
}

}  // namespace dlmcq

using namespace dlmcq;

extern "C" size_t dlmcq_host_staging_bytes(int64_t chunk_elems, int dtype) {
  if (chunk_elems < 1) return 0;
  const size_t es = dtype == DLMCQ_F32 ? 4 : 2;
  const size_t slot = 4 * up(static_cast<size_t>(chunk_elems) * es) + up(dlmcq_workspace_bytes(nullptr)) + up(64);
  return kSlots * slot + up(static_cast<size_t>(kChunkRing) * sizeof(float)) +
         up(static_cast<size_t>(kCallRing) * sizeof(float));
}

extern "C" int dlmcq_host_synchronize(void) {
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  StreamPool* pool = get_pool();
  if (!pool) return set_cuda_error(cudaGetLastError());
  return sync_pool(pool);
}

extern "C" int dlmcq_host_fq_forward_backward_async(const void* x_host, const void* dy_host, void* y_host,
                                                    void* dx_host, float* dscale_host, int64_t numel, int dtype,
                                                    int form, int lo, int hi, float g, float scale, float offset,
                                                    void* device_staging, size_t staging_bytes, int64_t chunk_elems) {
  if (!x_host || !dy_host || !y_host || !dx_host || !dscale_host || !device_staging || numel < 0 || chunk_elems < 1)
    return DLMCQ_EINVAL;
  if (dtype != DLMCQ_F32 && dtype != DLMCQ_BF16) return DLMCQ_EINVAL;
  if (chunk_elems % 8 != 0) return DLMCQ_EINVAL;   // keeps every chunk 16-byte aligned
  if (staging_bytes < dlmcq_host_staging_bytes(chunk_elems, dtype)) return DLMCQ_EWORKSPACE;
  const int64_t nchunks = (numel + chunk_elems - 1) / chunk_elems;
  if (nchunks > kChunkRing / 2) return DLMCQ_EUNSUPPORTED;
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  StreamPool* pool = get_pool();
  if (!pool) return set_cuda_error(cudaGetLastError());
  // ring hygiene: never let un-synchronised calls wrap onto entries that may still be read
  if (pool->in_flight_chunks + nchunks > kChunkRing / 2 || pool->call_cursor % kCallRing == kCallRing - 1) {
    if (int st = sync_pool(pool)) return st;
  }

  const size_t es = dtype == DLMCQ_F32 ? 4 : 2;
  const size_t buf = up(static_cast<size_t>(chunk_elems) * es);
  const size_t wsb = up(dlmcq_workspace_bytes(nullptr));
  const size_t slot = 4 * buf + wsb + up(64);
  char* base = static_cast<char*>(device_staging);
  float* d_chunk = reinterpret_cast<float*>(base + kSlots * slot);
  float* d_call = reinterpret_cast<float*>(base + kSlots * slot + up(static_cast<size_t>(kChunkRing) * sizeof(float)));

  cudaError_t e;
  const float params[2] = {scale, offset};
  const int64_t start = pool->chunk_cursor;
  bool used[kSlots] = {false};
  int status = DLMCQ_OK;
  for (int64_t c = 0; c < nchunks && status == DLMCQ_OK; ++c) {
    const int s = static_cast<int>(pool->rr++ % kSlots);
    cudaStream_t st = pool->s[s];
    char* sb = base + s * slot;
    void *dxin = sb, *ddy = sb + buf, *dyout = sb + 2 * buf, *ddx = sb + 3 * buf, *ws = sb + 4 * buf;
    float* d_params = reinterpret_cast<float*>(sb + 4 * buf + wsb);
    if (!used[s]) {
      // this call's qparams, ordered after the slot's earlier kernels by the stream itself
      if ((e = cudaMemcpyAsync(d_params, params, sizeof(params), cudaMemcpyHostToDevice, st)) != cudaSuccess) {
        status = set_cuda_error(e);
        break;
      }
      used[s] = true;
    }
    dlmcq_qparams qp;
    qp.form = form; qp.lo = lo; qp.hi = hi; qp.g = g; qp.scale = d_params; qp.offset = d_params + 1;
    const int64_t off = c * chunk_elems;
    const int64_t len = (numel - off) < chunk_elems ? (numel - off) : chunk_elems;
    const size_t bytes = static_cast<size_t>(len) * es;
    const char* xh = static_cast<const char*>(x_host) + off * es;
    const char* gh = static_cast<const char*>(dy_host) + off * es;
    if ((e = cudaMemcpyAsync(dxin, xh, bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
    if ((e = cudaMemcpyAsync(ddy, gh, bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
    dlmcq_layout l = {1, 1, len, dtype};
    status = dlmcq_fq_forward(dxin, dyout, nullptr, &l, &qp, st);
    if (status != DLMCQ_OK) break;
    if ((e = cudaMemcpyAsync(static_cast<char*>(y_host) + off * es, dyout, bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
    status = dlmcq_fq_backward(dxin, ddy, ddx, d_chunk + (start + c) % kChunkRing, nullptr, &l, &qp, ws, wsb, st);
    if (status != DLMCQ_OK) break;
    if ((e = cudaMemcpyAsync(static_cast<char*>(dx_host) + off * es, ddx, bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
  }
  if (status != DLMCQ_OK) return status;
  pool->chunk_cursor = (start + nchunks) % kChunkRing;
  pool->in_flight_chunks += nchunks;
  // epilogue: wait for this call's chunks on every slot, reduce, ship the scalar home
  for (int s = 0; s < kSlots; ++s) {
    if (!used[s]) continue;
    cudaEventRecord(pool->slot_done[s], pool->s[s]);
    cudaStreamWaitEvent(pool->epi, pool->slot_done[s], 0);
  }
  float* out = d_call + (pool->call_cursor++ % kCallRing);
  sum_chunks_kernel<<<1, 1, 0, pool->epi>>>(d_chunk, start, nchunks, kChunkRing, out);
  DLMCQ_LAUNCH_CHECK();
  if ((e = cudaMemcpyAsync(dscale_host, out, sizeof(float), cudaMemcpyDeviceToHost, pool->epi)) != cudaSuccess)
    return set_cuda_error(e);
  return DLMCQ_OK;
}

extern "C" int dlmcq_host_fq_forward_backward(const void* x_host, const void* dy_host, void* y_host, void* dx_host,
                                              float* dscale_host, int64_t numel, int dtype, int form, int lo, int hi,
                                              float g, float scale, float offset, void* device_staging,
                                              size_t staging_bytes, int64_t chunk_elems) {
  int st = dlmcq_host_fq_forward_backward_async(x_host, dy_host, y_host, dx_host, dscale_host, numel, dtype, form, lo,
                                                hi, g, scale, offset, device_staging, staging_bytes, chunk_elems);
  if (st != DLMCQ_OK) return st;
  return dlmcq_host_synchronize();
}
