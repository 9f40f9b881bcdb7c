// host_pipeline.cu - end-to-end entry for HOST-resident tensors.
//
// A caller that owns (pinned) host buffers - the situation of a plugin bound over ctypes/cgo/JNI
// with no device tensors of its own - gets fake-quant forward + backward in one call: the tensor
// is cut into chunks and each chunk flows H2D -> forward kernel -> backward kernel -> D2H on one
// of kSlots streams, so copies in both directions overlap the kernels.  The arithmetic is the
// same dlmcq_fq_forward / dlmcq_fq_backward code path (per-tensor qparams).
#include <mutex>

#include "common.cuh"

namespace dlmcq {

constexpr int kSlots = 3;
constexpr size_t kAlign = 256;
constexpr int64_t kMaxChunks = 65536;

static inline size_t up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct StreamPool {
  cudaStream_t s[kSlots];
  cudaEvent_t params_ready;
  bool ok = false;
};
static StreamPool g_pools[64];
static std::mutex g_pool_mutex;

static StreamPool* get_pool() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  StreamPool& p = g_pools[dev];
  if (!p.ok) {
    for (int i = 0; i < kSlots; ++i)
      if (cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&p.params_ready, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    p.ok = true;
  }
  return &p;
}

}  // namespace dlmcq

using namespace dlmcq;

extern "C" size_t dlmcq_host_staging_bytes(int64_t chunk_elems, int dtype) {
  if (chunk_elems < 1) return 0;
  const size_t es = dtype == DLMCQ_F32 ? 4 : 2;
  const size_t slot = 4 * up(static_cast<size_t>(chunk_elems) * es) + up(dlmcq_workspace_bytes(nullptr));
  return kSlots * slot + up(64) + up(static_cast<size_t>(kMaxChunks) * sizeof(float));
}

extern "C" int dlmcq_host_fq_forward_backward(const void* x_host, const void* dy_host, void* y_host, void* dx_host,
                                              float* dscale_host, int64_t numel, int dtype, int form, int lo, int hi,
                                              float g, float scale, float offset, void* device_staging,
                                              size_t staging_bytes, int64_t chunk_elems) {
  if (!x_host || !dy_host || !y_host || !dx_host || !dscale_host || !device_staging || numel < 0 || chunk_elems < 1)
    return DLMCQ_EINVAL;
  if (dtype != DLMCQ_F32 && dtype != DLMCQ_BF16) return DLMCQ_EINVAL;
  if (chunk_elems % 8 != 0) return DLMCQ_EINVAL;   // keeps every chunk 16-byte aligned
  if (staging_bytes < dlmcq_host_staging_bytes(chunk_elems, dtype)) return DLMCQ_EWORKSPACE;
  const int64_t nchunks = (numel + chunk_elems - 1) / chunk_elems;
  if (nchunks > kMaxChunks) return DLMCQ_EUNSUPPORTED;
  StreamPool* pool = get_pool();
  if (!pool) return set_cuda_error(cudaGetLastError());

  const size_t es = dtype == DLMCQ_F32 ? 4 : 2;
  const size_t buf = up(static_cast<size_t>(chunk_elems) * es);
  const size_t wsb = up(dlmcq_workspace_bytes(nullptr));
  const size_t slot = 4 * buf + wsb;
  char* base = static_cast<char*>(device_staging);
  float* d_params = reinterpret_cast<float*>(base + kSlots * slot);
  float* d_dscale = reinterpret_cast<float*>(base + kSlots * slot + up(64));

  cudaError_t e;
  const float params[2] = {scale, offset};
  if ((e = cudaMemcpyAsync(d_params, params, sizeof(params), cudaMemcpyHostToDevice, pool->s[0])) != cudaSuccess)
    return set_cuda_error(e);
  for (int s = 0; s < kSlots; ++s)
    if ((e = cudaMemsetAsync(base + s * slot + 4 * buf, 0, kWsHeaderBytes, pool->s[s])) != cudaSuccess)
      return set_cuda_error(e);
  cudaEventRecord(pool->params_ready, pool->s[0]);
  for (int s = 1; s < kSlots; ++s) cudaStreamWaitEvent(pool->s[s], pool->params_ready, 0);

  dlmcq_qparams qp;
  qp.form = form; qp.lo = lo; qp.hi = hi; qp.g = g; qp.scale = d_params; qp.offset = d_params + 1;
  int status = DLMCQ_OK;
  for (int64_t c = 0; c < nchunks && status == DLMCQ_OK; ++c) {
    const int s = static_cast<int>(c % kSlots);
    cudaStream_t st = pool->s[s];
    char* sb = base + s * slot;
    void *dxin = sb, *ddy = sb + buf, *dyout = sb + 2 * buf, *ddx = sb + 3 * buf, *ws = sb + 4 * buf;
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
    status = dlmcq_fq_backward(dxin, ddy, ddx, d_dscale + c, nullptr, &l, &qp, ws, wsb, st);
    if (status != DLMCQ_OK) break;
    if ((e = cudaMemcpyAsync(static_cast<char*>(y_host) + off * es, dyout, bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
    if ((e = cudaMemcpyAsync(static_cast<char*>(dx_host) + off * es, ddx, bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { status = set_cuda_error(e); break; }
  }
  // per-chunk scale gradients come back last, after every slot stream has drained
  for (int s = 0; s < kSlots; ++s) {
    e = cudaStreamSynchronize(pool->s[s]);
    if (e != cudaSuccess && status == DLMCQ_OK) status = set_cuda_error(e);
  }
  if (status != DLMCQ_OK) return status;
  double total = 0.0;
  if (nchunks > 0) {
    float* tmp = new float[nchunks];
    e = cudaMemcpy(tmp, d_dscale, static_cast<size_t>(nchunks) * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
      for (int64_t c = 0; c < nchunks; ++c) total += static_cast<double>(tmp[c]);   // fixed order
    delete[] tmp;
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  dscale_host[0] = static_cast<float>(total);
  return DLMCQ_OK;
}
