// LSTM recurrence on the tcgen05 tensor cores (AVHOT_PREC_TC path of a3, src/encoders.py:54-61, :148-154).
//
// Per layer, after the time-batched input projection xp = x W_ih^T + b (gemm_tc.cu), one persistent CTA owns
// 128 clips for the whole sequence and every step computes
//     gates[128 clips, 1024] = h_{t-1}[128, 256] (fp16, shared memory) x W_hh^T (fp16, streamed by TMA) + xp_t
// in 8 chunks of 32 hidden units (N = 4 gates x 32 = 128 accumulator columns, double-buffered in TMEM):
//   warp 0     TMA producer: W_hh tiles (128 gate rows x 64 k, 128-byte swizzle) through a 6-stage ring; the
//              1024 x 256 weight matrix is re-read from L2 every step (it never fits next to the state);
//   warp 1     one lane issues tcgen05.mma (M 128 clips, N 128, K 16) x 16 per chunk; A is the fp16 hidden
//              state the epilogue wrote, ping-ponged between steps;
//   warps 2-17 sixteen epilogue warps = (TMEM lane quarter) x (8-unit quarter of a chunk), thread = (clip, 8 units):
//              tcgen05.ld the four gate pre-activations of its units, add xp_t (coalesced loads, prefetched two
//              chunks ahead in registers), sigmoid / tanh, update the fp32 cell state that lives in 256 further
//              TMEM columns of the same lane (tcgen05.ld / .st), write h_t as fp16 straight into the swizzled
//              A-operand layout of the next step and (layer 0) into the next layer's fp16 GEMM operand.
// mbarriers: full/empty per weight stage, acc_full/acc_empty per TMEM buffer, h_ready per 64-unit k-block of
// the hidden state (so step t+1 starts on k-block 0 while the epilogue of step t finishes k-block 3).
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "tc.cuh"

namespace avhot {
namespace tc {

constexpr int RB = 128, RH = 256, RCH = 32, RNC = RH / RCH, RN = 4 * RCH, RKB = RH / 64, RST = 6, RTHREADS = 64 + 512 + 32;

struct RecSmem {
  __half hA[2][RKB][RB * 64];   // hidden state, A operand: [step parity][k-block][128 rows x 128 B, swizzled]
  __half wB[RST][RN * 64];      // W_hh tile ring, B operand
  uint64_t full[2 * RST], empty[2 * RST], acc_full[2], acc_empty[2], h_ready[RKB];
  uint64_t h_done[2][RKB], h_free[2][RKB];   // layer-output store: k-block of buffer p written / read out by the TMA store
  uint32_t tmem_base;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg256(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
// The pre-activations of the sigmoid gates arrive HALVED: W_hh, W_ih and the bias rows of i, f, o are packed times 0.5
// (exact), so act_sigmoid takes xh = x / 2.
// ACT = 0: 1 / (1 + 2^(-x log2 e)) by ex2.approx + rcp.approx (~2 ulp each), tanh = 2 sigmoid(2x) - 1: two MUFU
//          operations per activation.
// ACT = 1: tanh.approx.f32 (one MUFU, max relative error 2^-11), sigmoid(x) = 0.5 + 0.5 tanh(x / 2).
template <int ACT> __device__ __forceinline__ float act_sigmoid(float xh) {
  if (ACT == 0) return __fdividef(1.f, 1.f + __expf(-2.f * xh));
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(xh));
  return fmaf(0.5f, y, 0.5f);
}
template <int ACT> __device__ __forceinline__ float act_tanh(float x) {
  if (ACT == 0) return fmaf(2.f, __fdividef(1.f, 1.f + __expf(-2.f * x)), -1.f);
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// TRACE (debug, AVHOT_REC_TRACE=<file>): CTA 0 records clock64() at eight points of every chunk of steps
// [TRACE_T0, TRACE_T0 + TRACE_NT) into `trace` ([step][chunk][event]); see lstm_recurrence_tc for the dump.
constexpr int TRACE_T0 = 100, TRACE_NT = 4, TRACE_EV = 12;
// PAIR: two CTAs (a cluster of 2 = one tcgen05 CTA pair) share every W_hh tile.  The clock traces (r2c - r2e) and the
// half-grid experiment (r2j) showed the kernel paced per SM by the arrival of its operands - 512 KB of weights plus
// 256 KB of xp per step, ~40 B/clk/SM - not by the tensor pipe or the epilogue, and TMA multicast did not help (each SM
// still ingests every tile).  With cta_group::2 each CTA loads only 64 of a tile's 128 gate rows (8 KB) and the
// M = 256 MMA, issued by the leader CTA, reads both halves: the weight ingest per SM halves.  The leader's MMA thread
// owns the pipeline: both CTAs' TMA loads complete the leader's full[] barriers, both CTAs' epilogue warps arrive on
// the leader's acc_empty[] / h_ready[], and the leader's commits are multicast to both CTAs' empty[] / acc_full[].
// The grid is rounded up to an even size; a CTA past n_ct redoes the last clip tile and writes nothing.
template <int ACT, bool XP16, bool TRACE = false, bool PAIR = false>
__global__ void __launch_bounds__(RTHREADS, 1)
lstm_rec_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH,
                   const float* __restrict__ xp, __half* __restrict__ h_seq, int h_pitch,
                   float* __restrict__ h_last, const int32_t* __restrict__ lengths, int B, int T, int n_ct_arg,
                   long long* __restrict__ trace = nullptr) {
  const int ctile = min((int)blockIdx.x, n_ct_arg - 1);
  const bool writer = (int)blockIdx.x < n_ct_arg;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  constexpr int WROWS = PAIR ? RN / 2 : RN;          // gate rows of a weight tile held by this CTA
  // PAIR: a ring stage holds the four k-blocks of a whole chunk (4 x 8 KB per CTA), three stages in the same bytes:
  // the MMA thread, whose serial instruction stream paces the tensor pipe (the issue of tcgen05.mma blocks for about
  // the ~89 cycles an M128 N128 K16 takes, so every other instruction it executes between two MMAs is a pipe bubble:
  // r2p trace, 583 cycles per k-block of which 356 are MMAs), then waits and commits once per chunk, not per k-block
  constexpr int NST = PAIR ? 3 : RST;
  auto tr = [&](int t, int c, int ev) {
    if (TRACE && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && t >= TRACE_T0 && t < TRACE_T0 + TRACE_NT)
      trace[((t - TRACE_T0) * RNC + c) * TRACE_EV + ev] = clock64();
  };
  extern __shared__ unsigned char rec_smem_raw[];
  RecSmem& sm = *reinterpret_cast<RecSmem*>(rec_smem_raw + ((1024u - (smem_u32(rec_smem_raw) & 1023u)) & 1023u));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform

  // h_{-1} = 0
  {
    uint4* z = reinterpret_cast<uint4*>(&sm.hA[0][0][0]);
    for (int i = threadIdx.x; i < (int)(sizeof(sm.hA[0]) / 16); i += RTHREADS) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    constexpr int NC = PAIR ? 2 : 1;            // CTAs whose epilogue warps arrive on the leader's barriers
    for (int s = 0; s < NST; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sm.acc_full[b], 1); mbar_init(&sm.acc_empty[b], 16 * NC); }
    for (int k = 0; k < RKB; ++k) mbar_init(&sm.h_ready[k], 32 * NC);   // 16 epilogue warps x the 2 chunks of a k-block
    for (int k = 0; k < 2 * RKB; ++k) { mbar_init(&sm.h_done[0][0] + k, 32); mbar_init(&sm.h_free[0][0] + k, 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_pair(&sm.tmem_base, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(&sm.tmem_base, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();               // both CTAs' barriers are initialised before any remote arrival
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  __half* const ring = &sm.wB[0][0];          // PAIR: 2 RST stages of 8 KB in the same bytes

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      if (!PAIR) {
        for (int t = 0; t < T; ++t)
          for (int c = 0; c < RNC; ++c)
            for (int kb = 0; kb < RKB; ++kb, ++it) {
              const int s = it % NST;
              mbar_wait(&sm.empty[s], ((it / NST) & 1) ^ 1);
              mbar_expect_tx(&sm.full[s], (uint32_t)sizeof(sm.wB[0]));
              tma_load_2d(sm.wB[s], &tmW, &sm.full[s], kb * 64, c * RN);
            }
      } else {
        const uint32_t full0 = cluster_map_shared(&sm.full[0], 0);      // the leader's full[] barriers
        for (int t = 0; t < T; ++t)
          for (int c = 0; c < RNC; ++c, ++it) {
            const int s = it % NST;
            mbar_wait(&sm.empty[s], ((it / NST) & 1) ^ 1);
            tr(t, c, 8);
            if (leader) mbar_expect_tx(&sm.full[s], (uint32_t)(2 * RKB * WROWS * 64 * sizeof(__half)));   // 64 KB: both CTAs
            for (int kb = 0; kb < RKB; ++kb)
              tma_load_2d_pair(ring + (size_t)(s * RKB + kb) * (WROWS * 64), &tmW, full0 + (uint32_t)s * 8u, kb * 64,
                               c * RN + (int)crank * WROWS);
            tr(t, c, 9);
          }
      }
    }
  } else if (warp == 1) {
    if (leader) {                               // the whole warp, converged; one elected lane issues
      constexpr uint32_t idesc = umma_idesc(FMT_F16, PAIR ? 2 * RB : RB, RN);
      // An accumulator buffer is used once every two chunks and RNC / 2 = 4 times per step, so the parity of its use
      // count at chunk c is (c >> 1) & 1 in every step: no counters (as a run-time-indexed array they lived in local
      // memory, an LDL in front of every barrier wait of the thread that paces the kernel)
      static_assert((RNC / 2) % 2 == 0, "accumulator parity repeats every step");
      uint32_t it = 0;
      if (!PAIR) {
        for (int t = 0; t < T; ++t) {
          for (int c = 0; c < RNC; ++c) {
            const int b = c & 1;
            mbar_wait(&sm.acc_empty[b], (uint32_t)((c >> 1) & 1) ^ 1u);
            tc_fence_after();
            tr(t, c, 0);
            for (int kb = 0; kb < RKB; ++kb, ++it) {
              if (c == 0 && t > 0) mbar_wait(&sm.h_ready[kb], (uint32_t)(t - 1) & 1);
              if (kb == RKB - 1) tr(t, c, 7);
              const int s = it % NST;
              mbar_wait(&sm.full[s], (it / NST) & 1);
              tc_fence_after();
              if (kb == RKB - 1) tr(t, c, 1);
              const uint64_t adesc = umma_desc_sw128(smem_u32(sm.hA[t & 1][kb]));
              const uint64_t bdesc = umma_desc_sw128(smem_u32(sm.wB[s]));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_elect(tmem + (uint32_t)(b * RN), adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
              umma_commit_elect(&sm.empty[s]);
            }
            umma_commit_elect(&sm.acc_full[b]);
            tr(t, c, 2);
          }
        }
      } else {
        // the barriers of the next chunk are peeked (test_wait) while this chunk's last MMAs are still queued, so in
        // the steady state the thread goes from the last MMA of a chunk to the first of the next without a wait
        bool pre_acc = false, pre_full = false;
        for (int t = 0; t < T; ++t) {
          for (int c = 0; c < RNC; ++c, ++it) {
            const int b = c & 1, s = it % NST;
            if (TRACE && blockIdx.x == 0 && t >= TRACE_T0 && t < TRACE_T0 + TRACE_NT)
              trace[((t - TRACE_T0) * RNC + c) * TRACE_EV + 11] = (pre_acc ? 1 : 0) + (pre_full ? 2 : 0);
            tr(t, c, 10);
            if (!pre_acc) mbar_wait_cluster(&sm.acc_empty[b], (uint32_t)((c >> 1) & 1) ^ 1u);
            tr(t, c, 1);
            if (!pre_full) mbar_wait(&sm.full[s], (it / NST) & 1);
            tc_fence_after();
            tr(t, c, 0);
#pragma unroll
            for (int kb = 0; kb < RKB; ++kb) {
              if (c == 0 && t > 0) {
                mbar_wait_cluster(&sm.h_ready[kb], (uint32_t)(t - 1) & 1);
                tc_fence_after();
              }
              if (kb == RKB - 1) tr(t, c, 7);
              const uint64_t adesc = umma_desc_sw128(smem_u32(sm.hA[t & 1][kb]));
              const uint64_t bdesc = umma_desc_sw128(smem_u32(ring + (size_t)(s * RKB + kb) * (WROWS * 64)));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_pair_elect(tmem + (uint32_t)(b * RN), adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
              if (kb == RKB - 2) {
                const int bn = b ^ 1;
                pre_acc = mbar_test_wait_cluster(&sm.acc_empty[bn], (uint32_t)((((c + 1) % RNC) >> 1) & 1) ^ 1u);
                pre_full = mbar_test_wait_cluster(&sm.full[(it + 1) % NST], ((it + 1) / NST) & 1);
              }
            }
            umma_commit_pair_mc_elect(&sm.empty[s], 3);
            umma_commit_pair_mc_elect(&sm.acc_full[b], 3);
            tr(t, c, 2);
          }
        }
      }
    }
  } else if (warp >= 18) {
    // ---- store warp: the layer's output sequence (the next layer's GEMM operand) leaves through TMA.  A k-block of
    // h_t in the A-operand buffer is exactly a (128 clips x 64 units) box of the time-major (T * Bp, H) fp16 tensor in
    // the 128-byte-swizzled layout, so one bulk store per k-block and step replaces 512 scattered 16-byte stores per
    // chunk (the r2t trace: layer 0's epilogue tail, 300 - 1600 cycles per chunk, bound the whole kernel).
    if (h_seq && writer) {
      if (lane == 0) tma_prefetch_desc(&tmH);
      for (int t = 0; t < T; ++t) {
        const int pbuf = (t + 1) & 1;
        for (int kb = 0; kb < RKB; ++kb) {
          mbar_wait(&sm.h_done[pbuf][kb], (uint32_t)(t >> 1) & 1);      // all 16 epilogue warps wrote + fenced it
          if (lane == 0) {
            tma_store_2d(&tmH, sm.hA[pbuf][kb], kb * 64, (t * n_ct_arg + ctile) * RB);
            tma_store_commit();
            tma_store_wait_read0();
            mbar_arrive(&sm.h_free[pbuf][kb]);                            // the epilogue may overwrite it (step t + 2)
          }
          __syncwarp();
        }
      }
      if (lane == 0) tma_store_wait0();
    }
  } else {
    // 16 epilogue warps = (TMEM lane quarter, quarter q of a chunk): thread (clip r, q) owns hidden units
    // 32 c + 8 q .. + 7 of every chunk c, so the four 8-unit items of a chunk are computed in parallel by four
    // warps per lane quarter (4 epilogue warps per scheduler hide each other's TMEM / MUFU / memory latency) and a
    // chunk's epilogue takes one item time.  xp of the item two chunks ahead is in flight in registers.
    const int q = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // clip row = TMEM lane
    const int clip = ctile * RB + r;
    const bool live = writer && clip < B;
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
    const int last_t = (live && lengths ? min(max(__ldg(lengths + clip), 1), T) : T) - 1;
    {                                                         // c_{-1} = 0 for the units this thread owns
      uint32_t z[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = 0u;
#pragma unroll 1
      for (int c = 0; c < RNC; ++c) tmem_st8(tlane + 256u + (uint32_t)(c * RCH + q * 8), z);
      tmem_st_wait();
    }
    // xp is time-major and tile-transposed (gemm_ws.cu): [t][clip tile][4H columns][128 clips]; h_seq is the
    // time-major padded fp16 sequence (T, Bp, pitch) that is the next layer's GEMM operand (rows of dead clips are
    // written too, so the operand is fully initialised)
    const size_t n_ct = (size_t)n_ct_arg;
    const float* xcol = xp + (size_t)ctile * (4 * RH) * RB + r;
    const bool hstore = h_seq && writer;                     // this layer's h_t sequence goes out through the store warp
    unsigned char* const hbase = reinterpret_cast<unsigned char*>(&sm.hA[0][0][0]);

    // xp registers hold what was loaded, unconverted: the loads of the item two chunks ahead stay in flight until
    // the values are used (converting at load time made every thread wait for its own prefetch: a third of all
    // stall samples in the r1y profile).  fp32 xp: 4 gates x 8 floats; fp16 xp: 4 gates x one 128-bit octet.
    constexpr int XR = XP16 ? 16 : 32;
    auto load_x = [&](int t, int c, uint32_t (&xr)[XR]) {     // coalesced across the 32 clips of the warp
      if (t >= T) {
#pragma unroll
        for (int i = 0; i < XR; ++i) xr[i] = 0u;
      } else if (!XP16) {
        const float* src = xcol + ((size_t)t * n_ct * (4 * RH) + c * RCH + q * 8) * RB;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int j = 0; j < 8; ++j) xr[g * 8 + j] = __float_as_uint(ldg_stream(src + (size_t)(g * RH + j) * RB));
      } else {
        // fp16 xp: [t][clip tile][4H / 8 octets][128 clips][8 halves] (gemm_ws.cu)
        const uint4* src = reinterpret_cast<const uint4*>(xp) + ((size_t)t * n_ct + ctile) * (4 * RH / 8) * RB + r +
                           (size_t)((c * RCH + q * 8) / 8) * RB;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4* p = src + (size_t)(g * RH / 8) * RB;
          asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(xr[4 * g]), "=r"(xr[4 * g + 1]), "=r"(xr[4 * g + 2]), "=r"(xr[4 * g + 3])
                       : "l"(p));
        }
      }
    };
    auto xval = [&](const uint32_t (&xr)[XR], int g, int j) -> float {
      if (!XP16) return __uint_as_float(xr[g * 8 + j]);
      const __half2 h2 = *reinterpret_cast<const __half2*>(&xr[4 * g + (j >> 1)]);
      return (j & 1) ? __high2float(h2) : __low2float(h2);
    };
    // PAIR: the MMA thread that waits on these lives in the leader CTA
    const uint32_t acc_empty_addr[2] = {PAIR ? cluster_map_shared(&sm.acc_empty[0], 0) : 0u,
                                        PAIR ? cluster_map_shared(&sm.acc_empty[1], 0) : 0u};
    const uint32_t h_ready_addr0 = PAIR ? cluster_map_shared(&sm.h_ready[0], 0) : 0u;
    // b = c & 1 is passed as a literal from the two call sites: as a run-time index it put use[] and acc_empty_addr[] in
    // local memory (LDL / STL in front of every barrier wait)
    auto run = [&](int t, int c, uint32_t (&xg)[XR], const int b) {
      const int u0 = c * RCH + q * 8;
      mbar_wait(&sm.acc_full[b], (uint32_t)((c >> 1) & 1));
      tc_fence_after();
      if (threadIdx.x == 64) tr(t, c, 3);
      uint32_t acc[4][8], cs[8];
      tmem_st_wait();                          // the cell-state store of this thread's previous item has landed
#pragma unroll
      for (int g = 0; g < 4; ++g) tmem_ld8(tlane + (uint32_t)(b * RN + g * RCH + q * 8), acc[g]);
      tmem_ld8(tlane + 256u + (uint32_t)u0, cs);
      tmem_ld_wait();
      if (threadIdx.x == 64) tr(t, c, 4);
      // the accumulator is in registers: hand the TMEM buffer back to the MMA warp before the activation math
      // (one arrival per warp: 512 per-thread arrivals on one mbarrier serialised into ~1000 cycles per chunk -
      // the r2c / r2d clock traces showed the MMA thread passing acc_empty 600-800 cycles after the loads were done)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(acc_empty_addr[b]);
        else mbar_arrive(&sm.acc_empty[b]);
      }
      float hn[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float ig = act_sigmoid<ACT>(__uint_as_float(acc[0][j]) + xval(xg, 0, j));
        const float fg = act_sigmoid<ACT>(__uint_as_float(acc[1][j]) + xval(xg, 1, j));
        const float gg = act_tanh<ACT>(__uint_as_float(acc[2][j]) + xval(xg, 2, j));
        const float og = act_sigmoid<ACT>(__uint_as_float(acc[3][j]) + xval(xg, 3, j));
        const float cn = fmaf(fg, __uint_as_float(cs[j]), ig * gg);
        cs[j] = __float_as_uint(cn);
        hn[j] = og * act_tanh<ACT>(cn);
      }
      if (threadIdx.x == 64) tr(t, c, 5);
      tmem_st8(tlane + 256u + (uint32_t)u0, cs);
      // h_t (fp16) -> A operand of step t+1: k-block u0 / 64, 16-byte chunk (u0 % 64) / 8, XOR-swizzled by the row
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2 v = __floats2half2_rn(hn[2 * j], hn[2 * j + 1]);
        pk[j] = *reinterpret_cast<const uint32_t*>(&v);
      }
      // (a k-block of this buffer is read out by the store warp's TMA store issued two steps ago)
      if (hstore && t >= 2) mbar_wait(&sm.h_free[(t + 1) & 1][c >> 1], (uint32_t)((t - 2) >> 1) & 1);
      unsigned char* tile = hbase + (size_t)((t + 1) & 1) * sizeof(sm.hA[0]) + (size_t)(u0 >> 6) * (RB * 128) +
                            (size_t)r * 128;
      *reinterpret_cast<uint4*>(tile + ((((u0 & 63) >> 3) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      if (live && h_last && t == last_t) stg256(h_last + (size_t)clip * RH + u0, hn);
      fence_proxy_async_smem();                // this thread's part of k-block c/2 of h_t is written
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(h_ready_addr0 + (uint32_t)(c >> 1) * 8u);
        else mbar_arrive(&sm.h_ready[c >> 1]);
        if (hstore) mbar_arrive(&sm.h_done[(t + 1) & 1][c >> 1]);
      }
      // next use of this register buffer: the same quarter of chunk c + 2 (or of the next step)
      load_x(c + 2 < RNC ? t : t + 1, c + 2 < RNC ? c + 2 : c + 2 - RNC, xg);
      if (threadIdx.x == 64) tr(t, c, 6);
    };
    uint32_t xa[XR], xb[XR];
    load_x(0, 0, xa);
    load_x(0, 1, xb);
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
#pragma unroll 1
      for (int c = 0; c < RNC; c += 2) {
        run(t, c, xa, 0);
        run(t, c + 1, xb, 1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();               // no CTA leaves while its peer can still signal its barriers
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

// out[(c*128 + g*32 + u) * 256 + k] = half(W_hh[(g*256 + c*32 + u) * 256 + k])
__global__ void pack_whh_fp16_kernel(const float* __restrict__ w, __half* __restrict__ out) {
  const int row = blockIdx.x;                 // packed row
  const int c = row >> 7, g = (row >> 5) & 3, u = row & 31;
  const float* src = w + (size_t)(g * RH + c * RCH + u) * RH;
  const float sc = g == 2 ? 1.f : 0.5f;     // sigmoid gates stored halved (see act_sigmoid)
  for (int k = threadIdx.x; k < RH; k += blockDim.x) out[(size_t)row * RH + k] = __float2half_rn(sc * src[k]);
}

}  // namespace tc

int lstm_pack_whh_fp16(const float* w_hh, void* out, cudaStream_t s) {
  tc::pack_whh_fp16_kernel<<<4 * tc::RH, 128, 0, s>>>(w_hh, reinterpret_cast<__half*>(out));
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

template <typename Kern, typename... Args>
static int rec_launch(Kern kern, int cl, int n_ct, int smem, cudaStream_t s, Args... args) {
  AVHOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ceil_div(n_ct, cl) * cl));
  cfg.blockDim = dim3(tc::RTHREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cl > 1 ? 1 : 0;
  AVHOT_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  return AVHOT_OK;
}

// can every CTA pair of this grid be co-resident (one CTA per SM; a second wave would double the run time)?
template <typename K2>
static bool rec_pairs_fit(K2 k2, int n_ct, int smem) {
  static const int forced = [] { const char* e = getenv("AVHOT_REC_PAIR"); return e ? atoi(e) : -1; }();
  if (forced == 0) return false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ceil_div(n_ct, 2) * 2));
  cfg.blockDim = dim3(tc::RTHREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (cudaOccupancyMaxActiveClusters(&n, k2, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
  return n * 2 >= ceil_div(n_ct, 2) * 2;
}

int lstm_recurrence_tc(const void* whh_packed_fp16, const void* xp_any, bool xp_f16, void* h_seq_f16, int h_pitch,
                       float* h_last, const int32_t* lengths, int B, int T, cudaStream_t s) {
  const float* xp = reinterpret_cast<const float*>(xp_any);
  __half* h_seq = reinterpret_cast<__half*>(h_seq_f16);
  using namespace tc;
  const int smem = (int)sizeof(RecSmem) + 1024;
  const int n_ct = ceil_div(B, RB);
  // tanh.approx (one MUFU per activation) measured as accurate as ex2 + rcp here (5e-4 vs the fp64 oracle: the
  // error is set by the fp16 state); AVHOT_LSTM_ACT=0 selects the ex2 + rcp form.
  static const int act = [] { const char* e = getenv("AVHOT_LSTM_ACT"); return e ? atoi(e) : 1; }();
  const bool fast = act == 1 && xp_f16;                       // the benchmarked configuration
  const int cl = fast && n_ct > 1 && rec_pairs_fit(lstm_rec_tc_kernel<1, true, false, true>, n_ct, smem) ? 2 : 1;
  CUtensorMap tmW;
  int rc = make_tensor_map_2d(&tmW, whh_packed_fp16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 4 * RH, RH, RH, RN / cl, 64);
  if (rc != AVHOT_OK) return rc;
  CUtensorMap tmH = tmW;                        // unused without an output sequence
  if (h_seq) {
    AVHOT_REQUIRE(h_pitch >= RH && (h_pitch & 7) == 0, AVHOT_ERR_INVALID, "lstm_recurrence_tc: bad h_seq pitch %d", h_pitch);
    rc = make_tensor_map_2d(&tmH, h_seq, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)T * n_ct * RB, RH, (uint64_t)h_pitch,
                            RB, 64);
    if (rc != AVHOT_OK) return rc;
  }
  static const char* trace_path = getenv("AVHOT_REC_TRACE");
  if (trace_path && fast && T >= TRACE_T0 + TRACE_NT) {     // debug: one traced launch, dumped as text
    long long* d_trace = nullptr;
    const size_t n = (size_t)TRACE_NT * RNC * TRACE_EV;
    AVHOT_CUDA(cudaMalloc(&d_trace, n * sizeof(long long)));
    AVHOT_CUDA(cudaMemsetAsync(d_trace, 0, n * sizeof(long long), s));
    if (cl == 2) rc = rec_launch(lstm_rec_tc_kernel<1, true, true, true>, 2, n_ct, smem, s, tmW, tmH, xp, h_seq, h_pitch, h_last, lengths, B, T, n_ct, d_trace);
    else rc = rec_launch(lstm_rec_tc_kernel<1, true, true, false>, 1, n_ct, smem, s, tmW, tmH, xp, h_seq, h_pitch, h_last, lengths, B, T, n_ct, d_trace);
    if (rc != AVHOT_OK) return rc;
    AVHOT_LAUNCH_CHECK();
    std::vector<long long> host(n);
    AVHOT_CUDA(cudaMemcpyAsync(host.data(), d_trace, n * sizeof(long long), cudaMemcpyDeviceToHost, s));
    AVHOT_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_trace);
    if (FILE* f = fopen(trace_path, "a")) {
      fprintf(f, "# B=%d T=%d cluster=%d  events: 0 mma:acc_empty 7 mma:h_ready(kb3) 1 mma:w_full(kb3) 2 mma:commit 3 epi:acc_full 4 epi:ld 5 epi:math 6 epi:end 8 tma:empty 9 tma:issued 10 mma:top 11 peeks\n", B, T, cl);
      for (int i = 0; i < TRACE_NT * RNC; ++i) {
        fprintf(f, "t=%d c=%d", TRACE_T0 + i / RNC, i % RNC);
        for (int e = 0; e < TRACE_EV; ++e) fprintf(f, " %lld", e == 11 ? host[i * TRACE_EV + e] : host[i * TRACE_EV + e] - host[0]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    return AVHOT_OK;
  }
  ProfScope prof("lstm_recurrence_tc", s);
  long long* no_trace = nullptr;
#define AVHOT_REC_LAUNCH(ACT_, X16_, PAIR_)                                                                         \
  rc = rec_launch(lstm_rec_tc_kernel<ACT_, X16_, false, PAIR_>, PAIR_ ? 2 : 1, n_ct, smem, s, tmW, tmH, xp, h_seq, \
                  h_pitch, h_last, lengths, B, T, n_ct, no_trace)
  if (fast && cl == 2) AVHOT_REC_LAUNCH(1, true, true);
  else if (fast) AVHOT_REC_LAUNCH(1, true, false);
  else if (act == 1) AVHOT_REC_LAUNCH(1, false, false);
  else if (xp_f16) AVHOT_REC_LAUNCH(0, true, false);
  else AVHOT_REC_LAUNCH(0, false, false);
#undef AVHOT_REC_LAUNCH
  if (rc != AVHOT_OK) return rc;
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

}  // namespace avhot
