/* abi_consumer.c - a plain-C program that uses libdlmcq.so the way a non-Python host (cgo / JNI / a C service)
 * would: CUDA runtime for memory, include/dlmcq.h for everything else, no torch, no Python.  It checks the device
 * results against the plain-C oracle (oracle/fq_oracle.c, linked as liboracle.so - TEST INFRASTRUCTURE) and prints
 * one PASS/FAIL line per check plus the measured bandwidth; exit status 0 iff every check passed.
 *
 *   built by __graft_entry__.build() into tests/_build/abi_consumer; run by tests/test_abi_consumer.py (-m gpu)
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dlmcq.h"

/* oracle/fq_oracle.c */
void orc_fq_forward(const float* x, float* y, float* codes, int64_t n, int64_t channels, int64_t inner,
                    const float* scale, const float* offset, int form, float lo, float hi, float g);
void orc_fq_backward(const float* x, const float* dy, float* dx, double* dscale, int64_t n, int64_t channels,
                     int64_t inner, const float* scale, const float* offset, int form, float lo, float hi, float g);
void orc_minmax(const float* x, int64_t channels, int64_t inner, int n_bits, int is_signed, float* scale, float* offset);
float orc_grad_scale_value(float s, float g);
void orc_code_gemm(const float* a_codes, const float* w_codes, int64_t m, int64_t n, int64_t k, float m_a, float o_a,
                   float z_a, const float* m_w, int64_t m_w_count, const float* bias, int relu, float* out);

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } \
  } while (0)
#define QK(call)                                                                                   \
  do {                                                                                             \
    int s_ = (call);                                                                               \
    if (s_ != DLMCQ_OK) { fprintf(stderr, "dlmcq error %d (%s: %s) at %s:%d\n", s_, dlmcq_status_string(s_), dlmcq_last_cuda_error(), __FILE__, __LINE__); exit(3); } \
  } while (0)

static int failures = 0;
static void report(const char* what, int ok) {
  printf("%s %s\n", ok ? "PASS" : "FAIL", what);
  if (!ok) ++failures;
}

static uint64_t rng = 0x2333ULL;
static float urand(void) { /* xorshift64*, uniform in [0,1) */
  rng ^= rng >> 12; rng ^= rng << 25; rng ^= rng >> 27;
  return (float)((rng * 0x2545F4914F6CDD1DULL) >> 40) / 16777216.0f;
}
static float nrand(void) { return sqrtf(-2.f * logf(urand() + 1e-12f)) * cosf(6.2831853f * urand()); }

static float* dev_copy(const float* h, size_t n) {
  float* d;
  CK(cudaMalloc((void**)&d, n * sizeof(float)));
  CK(cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}

/* one fake-quant forward + backward case against the oracle */
static void fq_case(const char* name, int form, int lo, int hi, int64_t outer, int64_t channels, int64_t inner, int relu) {
  const int64_t n = outer * channels * inner;
  float *x = malloc(n * 4), *dy = malloc(n * 4), *y = malloc(n * 4), *dx = malloc(n * 4), *yr = malloc(n * 4), *dxr = malloc(n * 4);
  float *scale = malloc(channels * 4), *offset = malloc(channels * 4), *ds = malloc(channels * 4);
  double* dsr = malloc(channels * 8);
  for (int64_t i = 0; i < n; ++i) {
    const float v = nrand() * (relu ? 1.5f : 0.05f);
    x[i] = relu ? (v > 0.f ? v : 0.f) : v;
    dy[i] = nrand();
  }
  x[0] = -0.0f; x[1] = NAN; x[2] = INFINITY; x[3] = 1e-45f;                  /* special values travel too */
  for (int64_t c = 0; c < channels; ++c) {
    scale[c] = relu ? 0.21f + 0.01f * (float)c : 0.013f + 0.0005f * (float)c;
    offset[c] = (form == DLMCQ_FORM_ZP) ? 3.f : 0.f;
  }
  const float g = (form == DLMCQ_FORM_AFFINE) ? (float)(1.0 / sqrt((double)n * hi)) : 0.f;
  dlmcq_layout lay = {outer, channels, inner, DLMCQ_F32};
  float *d_x = dev_copy(x, n), *d_dy = dev_copy(dy, n), *d_s = dev_copy(scale, channels), *d_o = dev_copy(offset, channels);
  float *d_y, *d_dx, *d_ds;
  void* d_ws;
  CK(cudaMalloc((void**)&d_y, n * 4)); CK(cudaMalloc((void**)&d_dx, n * 4)); CK(cudaMalloc((void**)&d_ds, channels * 4));
  const size_t wsb = dlmcq_workspace_bytes(&lay);
  CK(cudaMalloc(&d_ws, wsb)); CK(cudaMemset(d_ws, 0, wsb));                  /* zeroed once; the library keeps it zeroed */
  dlmcq_qparams qp = {form, lo, hi, g, d_s, (form == DLMCQ_FORM_SYM) ? NULL : d_o};
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  QK(dlmcq_fq_forward(d_x, d_y, NULL, &lay, &qp, st));
  QK(dlmcq_fq_backward(d_x, d_dy, d_dx, d_ds, NULL, &lay, &qp, d_ws, wsb, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaMemcpy(y, d_y, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(dx, d_dx, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ds, d_ds, channels * 4, cudaMemcpyDeviceToHost));
  const float* off_ref = (form == DLMCQ_FORM_SYM) ? NULL : offset;
  orc_fq_forward(x, yr, NULL, n, channels, inner, scale, off_ref, form, (float)lo, (float)hi, g);
  orc_fq_backward(x, dy, dxr, dsr, n, channels, inner, scale, off_ref, form, (float)lo, (float)hi, g);
  int64_t bad_y = 0, bad_dx = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int nan_a = y[i] != y[i], nan_b = yr[i] != yr[i];
    if (nan_a != nan_b || (!nan_a && memcmp(&y[i], &yr[i], 4) != 0)) ++bad_y;  /* bit-exact, all NaNs alike */
    if ((dx[i] == 0.f) != (dxr[i] == 0.f) || fabsf(dx[i] - dxr[i]) > 1e-6f * fabsf(dxr[i])) ++bad_dx;
  }
  int ok_ds = 1;
  for (int64_t c = 0; c < channels; ++c) {
    if (dsr[c] != dsr[c]) { ok_ds = ok_ds && (ds[c] != ds[c]); continue; }      /* NaN input row -> NaN gradient */
    double floor_ = 0.0;
    for (int64_t i = 0; i < n; ++i) if ((i / inner) % channels == c && dy[i] == dy[i]) floor_ += fabs(dy[i]);
    floor_ *= 1e-7 * (fabs((double)lo) > fabs((double)hi) ? fabs((double)lo) : fabs((double)hi)) * (g > 0.f ? g : 1.f);
    if (fabs((double)ds[c] - dsr[c]) > 1e-5 * fabs(dsr[c]) + floor_ + 1e-12) ok_ds = 0;
  }
  char msg[256];
  snprintf(msg, sizeof msg, "%s: forward bit-exact vs the C oracle (%lld elements)", name, (long long)n);
  report(msg, bad_y == 0);
  snprintf(msg, sizeof msg, "%s: dx mask identical, values within 1e-6", name);
  report(msg, bad_dx == 0);
  snprintf(msg, sizeof msg, "%s: dscale within 1e-5 relative (+ reduction-order floor)", name);
  report(msg, ok_ds);
  CK(cudaFree(d_x)); CK(cudaFree(d_dy)); CK(cudaFree(d_s)); CK(cudaFree(d_o)); CK(cudaFree(d_y)); CK(cudaFree(d_dx));
  CK(cudaFree(d_ds)); CK(cudaFree(d_ws)); CK(cudaStreamDestroy(st));
  free(x); free(dy); free(y); free(dx); free(yr); free(dxr); free(scale); free(offset); free(ds); free(dsr);
}

static void observer_case(void) {
  const int64_t c = 96, k = 577;
  float *w = malloc(c * k * 4), *s = malloc(c * 4), *o = malloc(c * 4), *sr = malloc(c * 4), *orf = malloc(c * 4);
  for (int64_t i = 0; i < c * k; ++i) w[i] = nrand() * 0.03f;
  dlmcq_layout lay = {1, c, k, DLMCQ_F32};
  float *d_w = dev_copy(w, c * k), *d_stats, *d_s, *d_o;
  void* d_ws;
  const size_t wsb = dlmcq_workspace_bytes(&lay);
  CK(cudaMalloc((void**)&d_stats, c * 16)); CK(cudaMalloc((void**)&d_s, c * 4)); CK(cudaMalloc((void**)&d_o, c * 4));
  CK(cudaMalloc(&d_ws, wsb)); CK(cudaMemset(d_ws, 0, wsb));
  QK(dlmcq_obs_stats(d_w, d_stats, &lay, 0, d_ws, wsb, NULL));
  QK(dlmcq_obs_minmax_finalize(d_stats, d_s, d_o, c, 4, 1, 1, NULL));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(s, d_s, c * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o, d_o, c * 4, cudaMemcpyDeviceToHost));
  orc_minmax(w, c, k, 4, 1, sr, orf);
  report("per-channel min/max observer: scales and offsets bit-exact vs the C oracle",
         memcmp(s, sr, c * 4) == 0 && memcmp(o, orf, c * 4) == 0);
  CK(cudaFree(d_w)); CK(cudaFree(d_stats)); CK(cudaFree(d_s)); CK(cudaFree(d_o)); CK(cudaFree(d_ws));
  free(w); free(s); free(o); free(sr); free(orf);
}

/* the quantised layer's product on the integer codes: x [m,k] (QBase A4, float offset) and w [n,k] (W4 symmetric per
 * output channel) -> one byte per code -> alpha / beta from the device-resident qparams -> TMA + tcgen05 GEMM,
 * against the plain-C oracle on the oracle's own codes: bit-exact */
static void qgemm_case(int encoding, const char* name) {
  const int64_t m = 777, n = 200, k = 192;
  float *x = malloc(m * k * 4), *w = malloc(n * k * 4), *ca = malloc(m * k * 4), *cw = malloc(n * k * 4);
  float *sw = malloc(n * 4), *bias = malloc(n * 4), *out = malloc(m * n * 4), *ref = malloc(m * n * 4);
  for (int64_t i = 0; i < m * k; ++i) { const float v = nrand() * 1.5f; x[i] = (v > 0.f ? v : 0.f) + 0.05f; }
  for (int64_t i = 0; i < n * k; ++i) w[i] = nrand() * 0.05f;
  for (int64_t j = 0; j < n; ++j) { sw[j] = 0.02f + 0.0003f * (float)j; bias[j] = nrand(); }
  const float sa = 0.31f, oa = 0.05f;
  const float g = (float)(1.0 / sqrt((double)(m * k) * 15));
  orc_fq_forward(x, NULL, ca, m * k, 1, m * k, &sa, &oa, DLMCQ_FORM_AFFINE, 0.f, 15.f, g);
  orc_fq_forward(w, NULL, cw, n * k, n, k, sw, NULL, DLMCQ_FORM_SYM, -7.f, 7.f, 0.f);
  orc_code_gemm(ca, cw, m, n, k, orc_grad_scale_value(sa, g), oa, 0.f, sw, n, bias, 1, ref);

  float *d_x = dev_copy(x, m * k), *d_w = dev_copy(w, n * k), *d_sa = dev_copy(&sa, 1), *d_oa = dev_copy(&oa, 1);
  float *d_sw = dev_copy(sw, n), *d_bias = dev_copy(bias, n), *d_alpha, *d_beta, *d_out;
  void *d_ca, *d_cw;
  CK(cudaMalloc(&d_ca, m * k)); CK(cudaMalloc(&d_cw, n * k));
  CK(cudaMalloc((void**)&d_alpha, n * 4)); CK(cudaMalloc((void**)&d_beta, n * 4)); CK(cudaMalloc((void**)&d_out, m * n * 4));
  dlmcq_layout la = {1, 1, m * k, DLMCQ_F32}, lw = {1, n, k, DLMCQ_F32};
  dlmcq_qparams qa = {DLMCQ_FORM_AFFINE, 0, 15, g, d_sa, d_oa}, qw = {DLMCQ_FORM_SYM, -7, 7, 0.f, d_sw, NULL};
  QK(dlmcq_codes_forward(d_x, d_ca, &la, &qa, encoding, NULL));
  QK(dlmcq_codes_forward(d_w, d_cw, &lw, &qw, encoding, NULL));
  QK(dlmcq_qgemm_prepare(d_cw, n, k, encoding, &qa, &qw, n, d_bias, d_alpha, d_beta, NULL));
  QK(dlmcq_qgemm(d_ca, d_cw, d_alpha, d_beta, d_out, m, n, k, encoding, 0, 1, DLMCQ_F32, NULL));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, d_out, m * n * 4, cudaMemcpyDeviceToHost));
  report(name, memcmp(out, ref, m * n * 4) == 0);
  CK(cudaFree(d_x)); CK(cudaFree(d_w)); CK(cudaFree(d_sa)); CK(cudaFree(d_oa)); CK(cudaFree(d_sw)); CK(cudaFree(d_bias));
  CK(cudaFree(d_ca)); CK(cudaFree(d_cw)); CK(cudaFree(d_alpha)); CK(cudaFree(d_beta)); CK(cudaFree(d_out));
  free(x); free(w); free(ca); free(cw); free(sw); free(bias); free(out); free(ref);
}

static void bandwidth(void) {
  const int64_t n = (int64_t)1 << 26;
  float *d_x, *d_dy, *d_y, *d_dx, *d_ds, *d_s, *d_o;
  void* d_ws;
  dlmcq_layout lay = {1, 1, n, DLMCQ_F32};
  const size_t wsb = dlmcq_workspace_bytes(&lay);
  CK(cudaMalloc((void**)&d_x, n * 4)); CK(cudaMalloc((void**)&d_dy, n * 4)); CK(cudaMalloc((void**)&d_y, n * 4));
  CK(cudaMalloc((void**)&d_dx, n * 4)); CK(cudaMalloc((void**)&d_ds, 4)); CK(cudaMalloc(&d_ws, wsb));
  CK(cudaMemset(d_ws, 0, wsb)); CK(cudaMemset(d_x, 0x3c, n * 4)); CK(cudaMemset(d_dy, 0x3b, n * 4));
  const float sc = 0.2f, of = 0.f;
  d_s = dev_copy(&sc, 1); d_o = dev_copy(&of, 1);
  dlmcq_qparams qp = {DLMCQ_FORM_AFFINE, 0, 15, 1e-4f, d_s, d_o};
  cudaEvent_t a, b, c;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); CK(cudaEventCreate(&c));
  for (int w = 0; w < 3; ++w) { QK(dlmcq_fq_forward(d_x, d_y, NULL, &lay, &qp, NULL)); QK(dlmcq_fq_backward(d_x, d_dy, d_dx, d_ds, NULL, &lay, &qp, d_ws, wsb, NULL)); }
  const int reps = 20;
  CK(cudaEventRecord(a, NULL));
  for (int r = 0; r < reps; ++r) QK(dlmcq_fq_forward(d_x, d_y, NULL, &lay, &qp, NULL));
  CK(cudaEventRecord(b, NULL));
  for (int r = 0; r < reps; ++r) QK(dlmcq_fq_backward(d_x, d_dy, d_dx, d_ds, NULL, &lay, &qp, d_ws, wsb, NULL));
  CK(cudaEventRecord(c, NULL));
  CK(cudaEventSynchronize(c));
  float tf, tb;
  CK(cudaEventElapsedTime(&tf, a, b)); CK(cudaEventElapsedTime(&tb, b, c));
  printf("INFO 2^26 fp32 elements through the C ABI: forward %.1f us = %.0f GB/s, backward %.1f us = %.0f GB/s\n",
         tf / reps * 1e3, 8.0 * n / (tf / reps * 1e-3) / 1e9, tb / reps * 1e3, 12.0 * n / (tb / reps * 1e-3) / 1e9);
  CK(cudaFree(d_x)); CK(cudaFree(d_dy)); CK(cudaFree(d_y)); CK(cudaFree(d_dx)); CK(cudaFree(d_ds)); CK(cudaFree(d_ws));
  CK(cudaFree(d_s)); CK(cudaFree(d_o));
}

int main(void) {
  if (dlmcq_version() != DLMCQ_VERSION) { fprintf(stderr, "header / library version mismatch\n"); return 4; }
  fq_case("QBase A4 per-tensor activation (AFFINE), ragged length", DLMCQ_FORM_AFFINE, 0, 15, 1, 1, (1 << 20) + 3, 1);
  fq_case("FSPTQ W4 per-channel weights (SYM) [64, 577]", DLMCQ_FORM_SYM, -7, 7, 1, 64, 577, 0);
  fq_case("FSPTQ A8 zero-point activation (ZP)", DLMCQ_FORM_ZP, 0, 255, 1, 1, 300007, 1);
  fq_case("QBase A4 per-channel activation [6, 8, 28, 28] (AFFINE, tiled kernels)", DLMCQ_FORM_AFFINE, 0, 15, 6, 8, 784, 1);
  fq_case("QBase A4 per-channel activation [40, 8, 7, 7] (AFFINE, channel-major kernels)", DLMCQ_FORM_AFFINE, 0, 15, 40, 8, 49, 1);
  observer_case();
  qgemm_case(DLMCQ_QGEMM_I8, "integer-code layer product (tcgen05 kind::i8) [777,192] x [200,192]: bit-exact vs the C oracle");
  qgemm_case(DLMCQ_QGEMM_E4M3, "integer-code layer product (tcgen05 kind::f8f6f4, e4m3 codes): bit-exact vs the C oracle");
  bandwidth();
  printf("%s: %d check(s) failed\n", failures ? "FAILED" : "ALL PASS", failures);
  return failures ? 1 : 0;
}
