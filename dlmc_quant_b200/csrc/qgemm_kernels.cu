// qgemm_kernels.cu - the layer's matrix product on the INTEGER CODES (consumer side).
//
// Reference: dlmc/quantization/scalar/modules/linear.py / conv.py `_forward_func(q_input, q_weight)` called from
// modules/base.py:140 - F.linear / F.conv2d on the two fake-quantised fp32 tensors.  For a per-tensor activation
// quantizer and a per-output-channel (or per-tensor) weight quantizer without a weight offset the product factors:
//
//     y_a[m,k] = (ca[m,k] - z_a) * m_a + o_a          (z_a: FORM_ZP zero-point, o_a: FORM_A1 / AFFINE offset)
//     y_w[n,k] =  cw[n,k] * m_w[n]
//     sum_k y_a*y_w = m_a*m_w[n] * sum_k ca*cw  +  (o_a - z_a*m_a) * m_w[n] * sum_k cw[n,k]
//                   = alpha[n] * acc[m,n] + beta[n]
//
// acc is an exact integer dot product of the codes.  It runs on the 5th-generation tensor cores:
//   * operands: one BYTE per code (1/4 of the fp32 activation traffic the fake-quantised tensor costs cuDNN),
//     K-major, staged by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) through a 4-stage mbarrier pipeline;
//   * tcgen05.mma cta_group::1 issued by ONE thread, M=128 x N=128 x K=32 per instruction, accumulators in TMEM
//     (128 lanes x 128 columns); kind::i8 (u8|s8 x s8 -> s32, exact for every code width up to 8 bits; sm_100a has
//     it) or kind::f8f6f4 on codes stored as e4m3 bytes (exact for |code| <= 16, fp32 accumulators);
//   * epilogue: tcgen05.ld (32x32b.x32) -> registers -> a padded smem transpose -> coalesced 128-byte row stores of
//     out = RN(RN(acc * alpha[n]) + beta[n]) (+ ReLU), fp32 or bf16.
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocation + MMA issue, warps 2-17 = epilogue in
// two groups of 8 (warp w may only touch TMEM lanes 32*(w%4) .. +31: two warps per lane quarter and group, 64 of the
// tile's 128 columns each).  Persistent, one CTA per SM (206 KB of smem, 256 TMEM columns = two accumulator
// buffers, one per epilogue group): while one group streams its tile out, the other reads the next accumulator and
// the loads and MMAs of the tiles after that are already running.  This path is bound by the bytes it moves (codes in, y out), not by the
// tensor pipe - the point of doing the product on codes is the 4x smaller operand.
//
// Also here: dlmcq_codes_forward (x -> one byte per code, the same arithmetic as dlmcq_fq_forward, streaming) and
// dlmcq_qgemm_prepare (alpha / beta from the DEVICE-resident qparams: no host sync, learnable scales stay put).
#include <cuda.h>       // CUtensorMap + enums only; the encoder is fetched through cudaGetDriverEntryPoint
#include <cuda_fp8.h>

#include "fq_math.cuh"

namespace dlmcq {
namespace {

constexpr int kBM = 128, kBN = 128, kBK = 128;                // tile; kBK bytes = codes (one byte each)
constexpr int kStages = 4;
constexpr int kABytes = kBM * kBK, kBBytes = kBN * kBK, kStageBytes = kABytes + kBBytes;
constexpr int kUmmaK = 32;                                    // codes one tcgen05.mma of an 8-bit kind consumes
constexpr int kEpiWarps = 16;                                 // two groups of 8 (two per TMEM lane quarter, 64 columns each)
constexpr int kEpiGroupWarps = kEpiWarps / 2;                 // group g drains accumulator buffer g: even / odd tiles
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
constexpr int kTmemCols = 2 * kBN;                            // double-buffered accumulator, one 32-bit column per output column
constexpr int kStgPitch = 36;                                 // floats per row of the epilogue's transpose buffer
constexpr int kStgBytes = kEpiWarps * 32 * kStgPitch * 4;
constexpr int kBarBytes = 128;
constexpr size_t kGemmSmem = 1024 + static_cast<size_t>(kStages) * kStageBytes + kStgBytes + kBarBytes;
constexpr long long kWatchdogCycles = 30000000000LL;          // ~15 s in ONE barrier wait: a wedged pipeline traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (spin & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWatchdogCycles) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05.commit: the mbarrier receives one arrival when every tcgen05.mma this thread issued so far has completed
// (it implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Shared-memory matrix descriptor of a K-major operand tile laid out by TMA with the 128-byte swizzle: rows of 128
// bytes, 8-row groups 1024 bytes apart (stride byte offset), descriptor version 1 (sm_100), layout type 2
// (SWIZZLE_128B); the leading byte offset is not used by swizzled K-major layouts (canonical value 1).  The tile
// base is 1024-byte aligned, so a K step of 32 bytes inside the swizzle atom is a plain advance of the start address.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if (KIND == DLMCQ_QGEMM_I8) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"          // same statement: the registers may not be read before the wait
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// out[m, n] = RN(RN(float(sum_k A[m,k] * B[n,k]) * alpha[n]) + beta[n])   (optionally max(., 0))
//
// Persistent: one CTA per SM walks the 128x128 output tiles (N tiles fastest, so the CTAs that share an A tile run
// at the same time and the second read comes from L2).  Three decoupled pipelines:
//   TMA producer  --full[s]/empty[s]-->  MMA thread  --acc_full[b]/acc_empty[b]-->  8 epilogue warps
// The accumulator is double-buffered in TMEM (2 x 128 columns), so the epilogue of tile i overlaps the loads and
// MMAs of tile i+1, and the operand ring keeps running across tile boundaries.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <bool IS_INT>
__device__ __forceinline__ float acc_value(uint32_t v) {
  return IS_INT ? static_cast<float>(static_cast<int>(v)) : __uint_as_float(v);
}

// Per-column epilogue constants of one 32-column piece, fetched BEFORE the warp waits for the accumulator (their
// latency hides behind the wait).  MODE 0: one column per lane; 1: four (fp32 line stores); 2: eight (bf16).
template <int MODE>
struct EpiConst {
  static constexpr int W = MODE == 0 ? 1 : (MODE == 1 ? 4 : 8);
  float al[W], be[W];
  __device__ __forceinline__ void load(const float* __restrict__ alpha, const float* __restrict__ beta, int col0, int lane,
                                       int N) {
    const int col = col0 + (MODE == 0 ? lane : (MODE == 1 ? (lane & 7) * 4 : (lane & 3) * 8));
#pragma unroll
    for (int e = 0; e < W; ++e) {
      const bool ok = col + e < N;
      al[e] = ok ? __ldg(alpha + col + e) : 0.f;
      be[e] = ok ? __ldg(beta + col + e) : 0.f;
    }
  }
};

// One 32-row x 32-column piece of the tile: this lane holds row `lane` (32 accumulators); the piece is transposed
// through a padded smem buffer (pitch 36 floats: 128-bit writes and reads are both conflict-free) so that global
// stores are whole 128-byte lines.  MODE 0: scalar stores (any N, any alignment); 1: fp32, 16-byte stores
// (N % 4 == 0); 2: bf16, 16-byte stores of 8 values (N % 8 == 0).  All smem reads are issued before the first
// store; `floor` is 0 for a fused ReLU and -inf otherwise (max.NaN: NaN propagates, -0 -> +0, like torch.relu).
template <bool IS_INT, int MODE>
__device__ __forceinline__ void epilogue_piece(const uint32_t (&v)[32], float* stg, int lane, int row0, int col0, int M,
                                               int N, const EpiConst<MODE>& k, void* __restrict__ out, float floor,
                                               int out_bf16) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + lane * kStgPitch + 4 * j) =
        make_float4(acc_value<IS_INT>(v[4 * j]), acc_value<IS_INT>(v[4 * j + 1]), acc_value<IS_INT>(v[4 * j + 2]),
                    acc_value<IS_INT>(v[4 * j + 3]));
  __syncwarp();
  const int rows = M - row0;                                    // >= 1; the common case is a full piece (>= 32)
  if (MODE == 1) {
    const int rsub = lane >> 3, c4 = (lane & 7) * 4, col = col0 + c4;
    if (col < N) {
      float* p = static_cast<float*>(out) + static_cast<size_t>(row0 + rsub) * N + col;
      const size_t step = static_cast<size_t>(4) * N;
#pragma unroll
      for (int h = 0; h < 2; ++h) {                             // two batches of four rows: 4 smem reads in flight, 16 live registers
        float4 t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = *reinterpret_cast<const float4*>(stg + ((h * 4 + i) * 4 + rsub) * kStgPitch + c4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 o;
          o.x = max_nan(__fadd_rn(__fmul_rn(t[i].x, k.al[0]), k.be[0]), floor);
          o.y = max_nan(__fadd_rn(__fmul_rn(t[i].y, k.al[1]), k.be[1]), floor);
          o.z = max_nan(__fadd_rn(__fmul_rn(t[i].z, k.al[2]), k.be[2]), floor);
          o.w = max_nan(__fadd_rn(__fmul_rn(t[i].w, k.al[3]), k.be[3]), floor);
          if ((h * 4 + i) * 4 + rsub < rows) st_stream(reinterpret_cast<float4*>(p + (h * 4 + i) * step), o);
        }
      }
    }
    __syncwarp();
  } else if (MODE == 2) {
    const int rsub = lane >> 2, c8 = (lane & 3) * 8, col = col0 + c8;
    if (col < N) {
      __nv_bfloat16* p = static_cast<__nv_bfloat16*>(out) + static_cast<size_t>(row0 + rsub) * N + col;
      const size_t step = static_cast<size_t>(8) * N;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 t[2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          t[i][0] = *reinterpret_cast<const float4*>(stg + ((h * 2 + i) * 8 + rsub) * kStgPitch + c8);
          t[i][1] = *reinterpret_cast<const float4*>(stg + ((h * 2 + i) * 8 + rsub) * kStgPitch + c8 + 4);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float o[8] = {t[i][0].x, t[i][0].y, t[i][0].z, t[i][0].w, t[i][1].x, t[i][1].y, t[i][1].z, t[i][1].w};
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = max_nan(__fadd_rn(__fmul_rn(o[e], k.al[e]), k.be[e]), floor);
          if ((h * 2 + i) * 8 + rsub < rows) st_stream(reinterpret_cast<uint4*>(p + (h * 2 + i) * step), Vec<__nv_bfloat16>::pack(o));
        }
      }
    }
    __syncwarp();
  } else {
    const int col = col0 + lane;
    if (col < N) {
      const int nr = rows < 32 ? rows : 32;
      for (int r = 0; r < nr; ++r) {
        const float o = max_nan(__fadd_rn(__fmul_rn(stg[r * kStgPitch + lane], k.al[0]), k.be[0]), floor);
        const size_t idx = static_cast<size_t>(row0 + r) * static_cast<size_t>(N) + static_cast<size_t>(col);
        if (out_bf16) static_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(o);
        else static_cast<float*>(out)[idx] = o;
      }
    }
    __syncwarp();
  }
}

// The tile loop of one epilogue warp.  The 16 epilogue warps form two groups: group g owns accumulator buffer g,
// i.e. the CTA's even / odd tiles, so while one group streams its tile to global memory the other one is already
// waiting for / reading the next accumulator - the store pipe never idles behind a barrier round trip (timeline:
// profiles/r02_qgemm_timeline_probe.log).  Inside a group a warp owns TMEM lane quarter q and 64 columns (two
// 32-column pieces, read one at a time: 32 live accumulator registers).  The buffer is handed back to the MMA thread
// by ONE arrival per warp, right after the warp's last TMEM read.
template <bool IS_INT, int MODE>
__device__ __forceinline__ void epilogue_loop(uint32_t tmem, uint32_t bar_acc_full, uint32_t bar_acc_empty, float* stg,
                                              int lane, int q, int grp, int chalf, int num_tiles, int tiles_n, int M,
                                              int N, const float* __restrict__ alpha, const float* __restrict__ beta,
                                              void* __restrict__ out, float floor, int out_bf16) {
  const uint32_t full = bar_acc_full + 8 * grp, empty = bar_acc_empty + 8 * grp;
  const uint32_t tbase = tmem + grp * kBN + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(chalf * 64);
  uint32_t ti = grp;
  for (int tile = blockIdx.x + grp * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, ti += 2) {
    const int m0 = (tile / tiles_n) * kBM, n0 = (tile % tiles_n) * kBN;
    const int col0 = n0 + chalf * 64, row0 = m0 + q * 32;
    const bool has0 = col0 < N, has1 = col0 + 32 < N;            // warp-uniform
    EpiConst<MODE> k0;
    k0.load(alpha, beta, col0, lane, N);
    mbar_wait(full, (ti >> 1) & 1);
    tc_fence_after();
    uint32_t v[32];
    if (has0) tmem_ld32(tbase, v);
    if (!has1) {                                                // nothing more to read from this buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(empty);
    }
    if (has0 && row0 < M) epilogue_piece<IS_INT, MODE>(v, stg, lane, row0, col0, M, N, k0, out, floor, out_bf16);
    if (has1) {
      EpiConst<MODE> k1;                                        // L1-resident by now; fetched here to keep registers free
      k1.load(alpha, beta, col0 + 32, lane, N);
      tmem_ld32(tbase + 32, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(empty);
      if (row0 < M) epilogue_piece<IS_INT, MODE>(v, stg, lane, row0, col0 + 32, M, N, k1, out, floor, out_bf16);
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(kGemmThreads, 1)   // 576 threads: ptxas caps the kernel at 96 registers
qgemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             const float* __restrict__ alpha, const float* __restrict__ beta, void* __restrict__ out, int M, int N,
             int K, uint32_t idesc, int relu, int out_bf16, int vec_ok) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // the swizzle atoms need 1024-byte alignment
  uint8_t* const tiles = smem_raw + (base - raw);
  float* const stg_all = reinterpret_cast<float*>(tiles + kStages * kStageBytes);
  const uint32_t bars = base + kStages * kStageBytes + kStgBytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kStages, bar_acc_full = bars + 16 * kStages,
                 bar_acc_empty = bar_acc_full + 16;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(tiles + kStages * kStageBytes + kStgBytes + 16 * kStages + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + kBN - 1) / kBN, tiles_m = (M + kBM - 1) / kBM;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);                           // one arrive.expect_tx by the producer
      mbar_init(bar_empty + 8 * s, 1);                          // one tcgen05.commit by the MMA thread
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full + 8 * b, 1);                       // one tcgen05.commit per tile
      mbar_init(bar_acc_empty + 8 * b, kEpiGroupWarps);         // one arrival per warp of the group that drains buffer b
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {                                              // the allocating warp also frees (whole warp, aligned)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * kBM, n0 = (tile % tiles_n) * kBN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);                 // slot free (passes at once in the first round)
          mbar_expect_tx(bar_full + 8 * s, kStageBytes);        // a box always delivers its full byte count (OOB = 0)
          const uint32_t dst = base + s * kStageBytes;
          tma_load_2d(dst, &map_a, bar_full + 8 * s, kb * kBK, m0);
          tma_load_2d(dst + kABytes, &map_b, bar_full + 8 * s, kb * kBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: one thread drives the tensor core for the whole CTA =====
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        const uint32_t ab = ti & 1, aph = (ti >> 1) & 1;
        mbar_wait(bar_acc_empty + 8 * ab, aph ^ 1);             // the epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d = tmem + ab * kBN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t a = base + s * kStageBytes, b = a + kABytes;
          const int left = K - kb * kBK;                        // bytes of K beyond this block are zero-filled: skip them
          const int steps = left >= kBK ? kBK / kUmmaK : (left + kUmmaK - 1) / kUmmaK;
          for (int k = 0; k < steps; ++k)
            umma<KIND>(d, umma_desc(a + k * kUmmaK), umma_desc(b + k * kUmmaK), idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit(bar_empty + 8 * s);                         // the slot returns to the producer when these complete
        }
        tc_commit(bar_acc_full + 8 * ab);                       // ... and this tile's accumulators are final
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> smem transpose -> 128-byte line stores =====
    constexpr bool IS_INT = KIND == DLMCQ_QGEMM_I8;
    const int ew = warp - 2;
    const int q = warp & 3;                                     // TMEM lane quarter this warp may read
    const int grp = ew >> 3;                                    // accumulator buffer / tile parity this warp serves
    const int chalf = (ew >> 2) & 1;                            // which 64 of the tile's 128 columns
    float* const stg = stg_all + ew * (32 * kStgPitch);
    const float floor = relu ? 0.f : __int_as_float(0xff800000);
    if (vec_ok && !out_bf16)
      epilogue_loop<IS_INT, 1>(tmem, bar_acc_full, bar_acc_empty, stg, lane, q, grp, chalf, num_tiles, tiles_n, M, N,
                               alpha, beta, out, floor, out_bf16);
    else if (vec_ok)
      epilogue_loop<IS_INT, 2>(tmem, bar_acc_full, bar_acc_empty, stg, lane, q, grp, chalf, num_tiles, tiles_n, M, N,
                               alpha, beta, out, floor, out_bf16);
    else
      epilogue_loop<IS_INT, 0>(tmem, bar_acc_full, bar_acc_empty, stg, lane, q, grp, chalf, num_tiles, tiles_n, M, N,
                               alpha, beta, out, floor, out_bf16);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// x -> one byte per code
// ---------------------------------------------------------------------------------------------------------------
template <int ENC>
__device__ __forceinline__ uint32_t code_byte(float code) {
  if (ENC == DLMCQ_QGEMM_E4M3)                                   // exact: the caller guarantees |code| <= 16
    return static_cast<uint32_t>(__nv_cvt_float_to_fp8(code == code ? code : 0.f, __NV_SATFINITE, __NV_E4M3));
  const int c = (code == code) ? static_cast<int>(code) : 0;     // NaN code -> 0 (as dlmcq_export_codes)
  return static_cast<uint32_t>(c) & 0xFFu;
}

// per-tensor, 16-byte aligned: every thread turns 16 elements into one 16-byte store (the arithmetic is fq_vec,
// i.e. bit-identical codes to dlmcq_fq_forward)
template <int FORM, typename T, int ENC>
__global__ void __launch_bounds__(kThreads, 4)
codes_flat_kernel(const T* __restrict__ x, uint8_t* __restrict__ out, int64_t n, const float* __restrict__ scale,
                  const float* __restrict__ offset, float g, float lo, float hi) {
  using V = Vec<T>;
  using raw = typename V::raw;
  constexpr int L = 16 / V::N;                                    // loads per 16 output bytes
  const ChanParams p = make_params<FORM>(scale, offset, 0, g, lo, hi);
  const int64_t groups = n / 16;
  const raw* xv = reinterpret_cast<const raw*>(x);
  uint4* ov = reinterpret_cast<uint4*>(out);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; gi < groups; gi += stride) {
    raw r[L];
#pragma unroll
    for (int k = 0; k < L; ++k) r[k] = ld_stream(xv + gi * L + k);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < L; ++k) {
      float f[V::N], fy[V::N], fc[V::N];
      V::unpack(r[k], f);
      fq_vec<FORM, V::N>(f, p, lo, hi, fc, fy);
#pragma unroll
      for (int e = 0; e < V::N; ++e) {
        const int b = k * V::N + e;
        w[b >> 2] |= code_byte<ENC>(fc[e]) << (8 * (b & 3));
      }
    }
    ov[gi] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (blockIdx.x == 0) {                                          // ragged tail (< 16 elements)
    const int64_t t = groups * 16 + threadIdx.x;
    if (threadIdx.x < 16 && t < n) {
      float c, v;
      fq_elem_ref<FORM>(to_f32<T>(x[t]), p, lo, hi, c, v);
      out[t] = static_cast<uint8_t>(code_byte<ENC>(c));
    }
  }
}

// any layout / alignment (per-channel weights: a one-off per weight update)
template <int FORM, typename T, int ENC>
__global__ void __launch_bounds__(kThreads)
codes_any_kernel(const T* __restrict__ x, uint8_t* __restrict__ out, int64_t n, int64_t channels, int64_t inner,
                 const float* __restrict__ scale, const float* __restrict__ offset, float g, float lo, float hi) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t ch = channels == 1 ? 0 : (i / inner) % channels;
    const ChanParams p = make_params<FORM>(scale, offset, ch, g, lo, hi);
    float c, v;
    fq_elem_ref<FORM>(to_f32<T>(x[i]), p, lo, hi, c, v);
    out[i] = static_cast<uint8_t>(code_byte<ENC>(c));
  }
}

template <int FORM, typename T, int ENC>
int codes_launch(const void* x, void* out, const dlmcq_layout* l, const dlmcq_qparams* qp, cudaStream_t st) {
  const int64_t n = l->outer * l->channels * l->inner;
  const float lo = static_cast<float>(qp->lo), hi = static_cast<float>(qp->hi);
  if (l->channels == 1 && aligned16(x) && aligned16(out)) {
    const int64_t tiles = (n / 16 + kThreads - 1) / kThreads;
    codes_flat_kernel<FORM, T, ENC><<<stream_grid(tiles, 16), kThreads, 0, st>>>(
        static_cast<const T*>(x), static_cast<uint8_t*>(out), n, qp->scale, qp->offset, qp->g, lo, hi);
  } else {
    const int64_t tiles = (n + kThreads - 1) / kThreads;
    codes_any_kernel<FORM, T, ENC><<<stream_grid(tiles, 8), kThreads, 0, st>>>(
        static_cast<const T*>(x), static_cast<uint8_t*>(out), n, l->channels, l->inner, qp->scale, qp->offset, qp->g,
        lo, hi);
  }
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

template <typename T, int ENC>
int codes_form(const void* x, void* out, const dlmcq_layout* l, const dlmcq_qparams* qp, cudaStream_t st) {
  switch (qp->form) {
    case DLMCQ_FORM_A1: return codes_launch<DLMCQ_FORM_A1, T, ENC>(x, out, l, qp, st);
    case DLMCQ_FORM_AFFINE: return codes_launch<DLMCQ_FORM_AFFINE, T, ENC>(x, out, l, qp, st);
    case DLMCQ_FORM_ZP: return codes_launch<DLMCQ_FORM_ZP, T, ENC>(x, out, l, qp, st);
    case DLMCQ_FORM_SYM: return codes_launch<DLMCQ_FORM_SYM, T, ENC>(x, out, l, qp, st);
  }
  return DLMCQ_EINVAL;
}

int code_range_ok(const dlmcq_qparams* qp, int encoding) {
  if (encoding == DLMCQ_QGEMM_E4M3) return (qp->lo >= -16 && qp->hi <= 16) ? DLMCQ_OK : DLMCQ_EUNSUPPORTED;
  if (encoding != DLMCQ_QGEMM_I8) return DLMCQ_EINVAL;
  if (qp->lo < -128 || qp->hi > 255 || (qp->lo < 0 && qp->hi > 127)) return DLMCQ_EUNSUPPORTED;
  return DLMCQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// alpha[n], beta[n] from the device-resident quantizer parameters; one warp per output channel
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ ChanParams params_of(int form, const float* scale, const float* offset, int64_t ch, float g,
                                                float lo, float hi) {
  switch (form) {
    case DLMCQ_FORM_A1: return make_params<DLMCQ_FORM_A1>(scale, offset, ch, g, lo, hi);
    case DLMCQ_FORM_AFFINE: return make_params<DLMCQ_FORM_AFFINE>(scale, offset, ch, g, lo, hi);
    case DLMCQ_FORM_ZP: return make_params<DLMCQ_FORM_ZP>(scale, offset, ch, g, lo, hi);
    default: return make_params<DLMCQ_FORM_SYM>(scale, offset, ch, g, lo, hi);
  }
}

__device__ __forceinline__ int decode_code(uint8_t b, int encoding, int is_signed) {
  if (encoding == DLMCQ_QGEMM_E4M3) {
    // small integers only: sign | exponent(4, bias 7) | mantissa(3)
    const int e = (b >> 3) & 0xF, m = b & 7;
    const float mag = e == 0 ? static_cast<float>(m) * 0.001953125f : ldexpf(1.f + static_cast<float>(m) * 0.125f, e - 7);
    const int v = static_cast<int>(mag);
    return (b & 0x80) ? -v : v;
  }
  return is_signed ? static_cast<int>(static_cast<int8_t>(b)) : static_cast<int>(b);
}

__global__ void __launch_bounds__(kRowWarps * 32)
qgemm_prepare_kernel(const uint8_t* __restrict__ w_codes, int64_t n, int64_t k, int encoding, int act_form,
                     const float* __restrict__ act_scale, const float* __restrict__ act_offset, float act_g,
                     float act_lo, float act_hi, int wt_form, const float* __restrict__ wt_scale, float wt_g,
                     float wt_lo, float wt_hi, int64_t wt_channels, const float* __restrict__ bias,
                     float* __restrict__ alpha, float* __restrict__ beta) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (row >= n) return;
  int sum = 0;
  const uint8_t* wr = w_codes + row * k;
  for (int64_t i = lane; i < k; i += 32) sum += decode_code(wr[i], encoding, wt_lo < 0.f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) {
    const ChanParams pa = params_of(act_form, act_scale, act_offset, 0, act_g, act_lo, act_hi);
    const ChanParams pw = params_of(wt_form, wt_scale, nullptr, wt_channels == 1 ? 0 : row, wt_g, wt_lo, wt_hi);
    const float z_a = act_form == DLMCQ_FORM_ZP ? pa.off : 0.f;
    const float o_a = (act_form == DLMCQ_FORM_A1 || act_form == DLMCQ_FORM_AFFINE) ? pa.off : 0.f;
    const float t = __fsub_rn(o_a, __fmul_rn(z_a, pa.mul));
    alpha[row] = __fmul_rn(pa.mul, pw.mul);
    float b = __fmul_rn(__fmul_rn(t, pw.mul), static_cast<float>(sum));
    if (bias) b = __fadd_rn(b, __ldg(bias + row));
    beta[row] = b;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static const EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// [rows, k] bytes, row-major (K-major): box = 128 bytes of K x 128 rows, 128-byte swizzle, out-of-bounds = 0
bool make_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t k, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(k)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// tcgen05 instruction descriptor: D format [4,6), A format [7,10), B format [10,13), A/B K-major (bits 15, 16 = 0),
// N >> 3 at [17,23), M >> 4 at [24,29)
uint32_t make_idesc(int encoding, int a_signed) {
  uint32_t d = 0;
  if (encoding == DLMCQ_QGEMM_I8) {
    d |= 2u << 4;                              // D = s32
    d |= (a_signed ? 1u : 0u) << 7;            // A = s8 | u8
    d |= 1u << 10;                             // B = s8
  } else {
    d |= 1u << 4;                              // D = f32; A = B = e4m3 (format 0)
  }
  d |= static_cast<uint32_t>(kBN >> 3) << 17;
  d |= static_cast<uint32_t>(kBM >> 4) << 24;
  return d;
}

template <int KIND>
int gemm_launch(const CUtensorMap& ma, const CUtensorMap& mb, const float* alpha, const float* beta, void* out, int m,
                int n, int k, int a_signed, int relu, int out_bf16, cudaStream_t st) {
  static bool opted_in[64] = {};                               // the attribute is per device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  if (dev < 0 || dev >= 64 || !opted_in[dev]) {
    const cudaError_t e = cudaFuncSetAttribute(qgemm_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(kGemmSmem));
    if (e != cudaSuccess) return set_cuda_error(e);
    if (dev >= 0 && dev < 64) opted_in[dev] = true;
  }
  const int64_t tiles = static_cast<int64_t>((n + kBN - 1) / kBN) * ((m + kBM - 1) / kBM);
  const int sms = num_sms();
  const unsigned grid = static_cast<unsigned>(tiles < sms ? tiles : sms);      // persistent: one CTA per SM
  // 16-byte row stores need every row start aligned: N a multiple of 4 (fp32) / 8 (bf16) and an aligned base
  const int vec_ok = aligned16(out) && (n % (out_bf16 ? 8 : 4) == 0);
  qgemm_kernel<KIND><<<grid, kGemmThreads, kGemmSmem, st>>>(ma, mb, alpha, beta, out, m, n, k,
                                                            make_idesc(KIND, a_signed), relu, out_bf16, vec_ok);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

}  // namespace
}  // namespace dlmcq

using namespace dlmcq;

extern "C" int dlmcq_codes_forward(const void* x, void* codes, const dlmcq_layout* layout, const dlmcq_qparams* qp,
                                   int encoding, void* stream) {
  if (!layout || !qp || !qp->scale || layout->outer < 1 || layout->channels < 1 || layout->inner < 0) return DLMCQ_EINVAL;
  if (layout->dtype != DLMCQ_F32 && layout->dtype != DLMCQ_BF16) return DLMCQ_EINVAL;
  if (int e = code_range_ok(qp, encoding)) return e;
  if (layout->outer * layout->channels * layout->inner == 0) return DLMCQ_OK;
  if (!x || !codes) return DLMCQ_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout->dtype == DLMCQ_F32)
    return encoding == DLMCQ_QGEMM_I8 ? codes_form<float, DLMCQ_QGEMM_I8>(x, codes, layout, qp, st)
                                      : codes_form<float, DLMCQ_QGEMM_E4M3>(x, codes, layout, qp, st);
  return encoding == DLMCQ_QGEMM_I8 ? codes_form<__nv_bfloat16, DLMCQ_QGEMM_I8>(x, codes, layout, qp, st)
                                    : codes_form<__nv_bfloat16, DLMCQ_QGEMM_E4M3>(x, codes, layout, qp, st);
}

extern "C" int dlmcq_qgemm_prepare(const void* w_codes, int64_t n, int64_t k, int encoding, const dlmcq_qparams* act_qp,
                                   const dlmcq_qparams* wt_qp, int64_t wt_channels, const float* bias, float* alpha,
                                   float* beta, void* stream) {
  if (!w_codes || !act_qp || !wt_qp || !act_qp->scale || !wt_qp->scale || !alpha || !beta || n < 1 || k < 1)
    return DLMCQ_EINVAL;
  if (wt_channels != 1 && wt_channels != n) return DLMCQ_EINVAL;
  if (int e = code_range_ok(act_qp, encoding)) return e;
  if (int e = code_range_ok(wt_qp, encoding)) return e;
  // the product only factors for weights without an additive term: y_w = code * m_w
  if (wt_qp->form == DLMCQ_FORM_ZP || wt_qp->offset != nullptr) return DLMCQ_EUNSUPPORTED;
  if (encoding == DLMCQ_QGEMM_I8 && wt_qp->lo >= 0 && wt_qp->hi > 127) return DLMCQ_EUNSUPPORTED;   // B operand is s8
  const int64_t blocks = (n + kRowWarps - 1) / kRowWarps;
  if (blocks > 0x7fffffffLL) return DLMCQ_EUNSUPPORTED;
  qgemm_prepare_kernel<<<static_cast<unsigned>(blocks), kRowWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(w_codes), n, k, encoding, act_qp->form, act_qp->scale, act_qp->offset, act_qp->g,
      static_cast<float>(act_qp->lo), static_cast<float>(act_qp->hi), wt_qp->form, wt_qp->scale, wt_qp->g,
      static_cast<float>(wt_qp->lo), static_cast<float>(wt_qp->hi), wt_channels, bias, alpha, beta);
  DLMCQ_LAUNCH_CHECK();
  return DLMCQ_OK;
}

extern "C" int dlmcq_qgemm(const void* a_codes, const void* w_codes, const float* alpha, const float* beta, void* out,
                           int64_t m, int64_t n, int64_t k, int encoding, int a_signed, int relu, int out_dtype,
                           void* stream) {
  if (m < 0 || n < 1 || k < 1 || !alpha || !beta) return DLMCQ_EINVAL;
  if (encoding != DLMCQ_QGEMM_I8 && encoding != DLMCQ_QGEMM_E4M3) return DLMCQ_EINVAL;
  if (out_dtype != DLMCQ_F32 && out_dtype != DLMCQ_BF16) return DLMCQ_EINVAL;
  if (m == 0) return DLMCQ_OK;
  if (!a_codes || !w_codes || !out) return DLMCQ_EINVAL;
  // TMA: 16-byte aligned base addresses and row pitch; 32-bit tile coordinates
  if (!aligned16(a_codes) || !aligned16(w_codes)) return DLMCQ_EALIGN;
  if ((k & 15) != 0 || m > 0x7fffff00LL || n > 0x7fffff00LL || k > 0x7fffff00LL) return DLMCQ_EUNSUPPORTED;
  if (((m + kBM - 1) / kBM) * ((n + kBN - 1) / kBN) > 0x7fffffffLL) return DLMCQ_EUNSUPPORTED;
  CUtensorMap ma, mb;
  if (!make_map(&ma, a_codes, m, k, kBM) || !make_map(&mb, w_codes, n, k, kBN)) return DLMCQ_EUNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int bf = out_dtype == DLMCQ_BF16;
  return encoding == DLMCQ_QGEMM_I8
             ? gemm_launch<DLMCQ_QGEMM_I8>(ma, mb, alpha, beta, out, static_cast<int>(m), static_cast<int>(n),
                                           static_cast<int>(k), a_signed, relu, bf, st)
             : gemm_launch<DLMCQ_QGEMM_E4M3>(ma, mb, alpha, beta, out, static_cast<int>(m), static_cast<int>(n),
                                             static_cast<int>(k), a_signed, relu, bf, st);
}
