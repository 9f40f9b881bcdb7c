// fold_kernels.cu - weight-space re-parameterisation feeding the per-channel observers.
//
//   mode DLMCQ_FOLD_MERGE_BN   dlmc/utils/merge_bn.py:84-100: fold a BatchNorm2d into the preceding Conv2d
//        var = running_var + 1e-7 ; w' = (w * gamma) / sqrt(var) ; b' = (gamma * (b - mean)) / sqrt(var) + beta
//   mode DLMCQ_FOLD_REPVGG     model/classification/repvgg.py:92-123: RepVGG block -> one 3x3 conv
//        std_k = sqrt(var_k + eps_k) ; t_k = gamma_k / std_k ; bias_k = beta_k - (mean_k * gamma_k) / std_k
//        w' = (k3 * t3 + pad(k1 * t1)) + id * t_id ;  b' = (bias3 + bias1) + bias_id
//
// The reference runs these as ~10 eager launches per layer and then re-reads the folded weights in the
// observer.  Here ALL layers of a model go through one launch (device descriptor table, one warp per
// (layer, output channel) row), and the same pass leaves every folded row's {min, max, max|w|, sum|w|}
// so that quantize_minmax_channel (ops.py:121-140) needs no second read.  Every reference op is one
// separately rounded fp32 instruction in the reference's order (additions of a literal 0 included: they
// turn -0.0 into +0.0); results are bit-identical.  Tiny tensors: the roofline is launch latency.
#include "common.cuh"

namespace dlmcq {

__device__ __forceinline__ int fold_find(const int64_t* __restrict__ prefix, int n_items, int64_t unit) {
  int lo = 0, hi = n_items;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(prefix + mid) <= unit) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kRowWarps * 32)
fold_grouped_kernel(const dlmcq_fold_item* __restrict__ items, const int64_t* __restrict__ chan_prefix, int n_items,
                    int64_t total_channels) {
  const int lane = threadIdx.x & 31;
  const int64_t gch = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (gch >= total_channels) return;
  const int k = fold_find(chan_prefix, n_items, gch);
  const dlmcq_fold_item it = items[k];
  const int64_t c = gch - __ldg(chan_prefix + k);
  const float* wr = it.w + c * it.inner;
  float* orow = it.w_out + c * it.inner;
  float mn = INFINITY, mx = -INFINITY, am = 0.f, sm = 0.f;
  auto stat = [&](float v) {
    const float a = fabsf(v);
    mn = fminf(mn, v); mx = fmaxf(mx, v); am = fmaxf(am, a); sm += a;
  };
  if (it.mode == DLMCQ_FOLD_MERGE_BN) {
    const float gamma = it.gamma[c], beta = it.beta[c], mean = it.mean[c];
    const float var = it.var[c] + 1e-7f;                                   // merge_bn.py:84
    const float sd = sqrtf(var);
    for (int64_t j = lane; j < it.inner; j += 32) {
      const float v = (wr[j] * gamma) / sd;                                // merge_bn.py:98
      orow[j] = v;
      stat(v);
    }
    if (lane == 0) {
      const float b = it.bias ? it.bias[c] : 0.f;                          // merge_bn.py:92-94: zeros if absent
      it.bias_out[c] = (gamma * (b - mean)) / sd + beta;                   // merge_bn.py:97
    }
  } else {
    // 3x3 branch
    const float std3 = sqrtf(it.var[c] + it.eps);                          // repvgg.py:121
    const float t3 = it.gamma[c] / std3;                                   // :122
    const float bias3 = it.beta[c] - (it.mean[c] * it.gamma[c]) / std3;    // :123
    // 1x1 branch (always present in a RepVGG block, repvgg.py:58)
    const float std1 = sqrtf(it.var1[c] + it.eps1);
    const float t1 = it.gamma1[c] / std1;
    const float bias1 = it.beta1[c] - (it.mean1[c] * it.gamma1[c]) / std1;
    // identity branch: BatchNorm only (repvgg.py:106-120), kernel = id_tensor with a 1 at [c, c % cin_g, 1, 1]
    const bool has_id = it.gamma_id != nullptr;
    float tid = 0.f, biasid = 0.f;
    if (has_id) {
      const float stdi = sqrtf(it.var_id[c] + it.eps_id);
      tid = it.gamma_id[c] / stdi;
      biasid = it.beta_id[c] - (it.mean_id[c] * it.gamma_id[c]) / stdi;
    }
    const int64_t kk = static_cast<int64_t>(it.ksize) * it.ksize;          // 9
    const int64_t centre = kk / 2;                                         // [1,1]
    const int64_t id_ci = c % it.cin_g;
    const float* w1r = it.w1 + c * it.cin_g;
    for (int64_t j = lane; j < it.inner; j += 32) {
      const int64_t ci = j / kk, p = j - ci * kk;
      const float a = wr[j] * t3;                                          // kernel3x3 * t
      const float b = (p == centre) ? w1r[ci] * t1 : 0.f;                  // pad(kernel1x1 * t, [1,1,1,1])
      // id_tensor * t: 1*t at the identity tap, 0*t elsewhere (a signed zero, or NaN for a non-finite t);
      // without the branch the reference adds the python int 0
      const float d = has_id ? (((p == centre) && (ci == id_ci)) ? 1.f : 0.f) * tid : 0.f;
      const float v = (a + b) + d;                                         // repvgg.py:96
      orow[j] = v;
      stat(v);
    }
    if (lane == 0) it.bias_out[c] = (bias3 + bias1) + biasid;              // repvgg.py:96
  }
  if (it.stats) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
      sm += __shfl_xor_sync(0xffffffffu, sm, o);
    }
    if (lane == 0) {
      const bool nan = sm != sm;                                           // torch.min / max propagate NaN
      float* s = it.stats + 4 * c;
      s[0] = nan ? sm : mn; s[1] = nan ? sm : mx; s[2] = nan ? sm : am; s[3] = sm;
    }
  }
}

}  // namespace dlmcq

using namespace dlmcq;

extern "C" int dlmcq_fold_grouped(const dlmcq_fold_item* items, const int64_t* chan_prefix, int n_items,
                                  int64_t total_channels, void* stream) {
  if (!items || !chan_prefix || n_items < 1 || total_channels < 0) return DLMCQ_EINVAL;
  if (total_channels == 0) return DLMCQ_OK;
  const int64_t blocks = (total_channels + kRowWarps - 1) / kRowWarps;
  if (blocks > 0x7fffffffLL) return DLMCQ_EUNSUPPORTED;
  fold_grouped_kernel<<<static_cast<unsigned>(blocks), kRowWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      items, chan_prefix, n_items, total_channels);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}
