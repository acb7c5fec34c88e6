// Weight-stationary tcgen05 GEMM for the LSTM input projections (AVHOT_PREC_TC path):
//     C[M, N] = A[M, K] (fp16) * W[N, K]^T (fp16) + bias[N] (fp32),   M = B*T ~ 10^6 rows, N = 1024, K <= 320.
// The op writes 2-4 KB per row against <= 640 B of A, and what bounded the first version (128-column blocks, ncu
// r1y / r2b: epilogue warps idle on acc_full half of the time, the MMA thread waiting for A tiles) was the L2 -> SM
// delivery of A: every SM pulled the whole A stream once per 128 output columns (~40 B/clk/SM, the same ceiling the
// recurrent kernel's weight stream hits; TMA multicast across the column-block CTAs did not move it - each SM still
// ingests every tile).  So a persistent CTA now owns a 256-column block of W for its whole life (160 KB of fp16 in
// shared memory, loaded once by TMA) and walks down M; A tiles (128 rows x 64 k, 128-byte swizzle) stream through a
// TMA ring (4 stages; 6 when K <= 256 leaves the fifth weight slot free - the ring depth is what covers the ~2 us
// loaded L2 latency) and are used for twice the columns; two 256-column TMEM accumulators alternate so the epilogue
// of tile i overlaps the MMAs of tile i+1.  K is consumed in steps of 16 (the last k-block may be partial).  The 4 CTAs that share an M tile walk the same tile sequence, so A is read
// from HBM once and served to the other three from L2.
//   warp 0     TMA producer (W block once, then the A ring)
//   warp 1     TMEM allocation + one lane issuing tcgen05.mma kind::f16 (M 128, N 256, K 16)
//   warps 2-9  two epilogue warpgroups (even / odd tiles): tcgen05.ld (thread = row) -> 128-bit global stores
// Layer 0's A is the fp16 3-way split [hi | lo | hi] of the features against W = [hi | hi | lo] (see lstm.cu), which
// makes this single fp16 pass fp32-class; layer 1's A is the fp16 hidden sequence written by the recurrent kernel.
#include <stdio.h>
#include <stdlib.h>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tc.cuh"

namespace avhot {
namespace tc {

constexpr int WBM = 128, WBN = 256, WKB_MAX = 5, WST = 4, WST_MAX = 6, WACC = 2, WTHREADS = 320;

struct GemmWsSmem {
  __half w[WKB_MAX][WBN * 64];   // the CTA's weight block, one 32 KB tile per 64-wide k-block
  __half a[WST][WBM * 64];       // A ring (stages 4, 5 live in w[4] when nkb <= 4)
  float bias[WBN];
  uint64_t w_full, full[WST_MAX], empty[WST_MAX], acc_full[WACC], acc_empty[WACC];
  uint32_t tmem_base;
};

template <bool OUT_F16>
__global__ void __launch_bounds__(WTHREADS, 1)
gemm_f16_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                   const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int nk16, int n_blocks,
                   int groups, long long* __restrict__ trace) {
  extern __shared__ unsigned char gemm_ws_smem_raw[];
  GemmWsSmem& sm = *reinterpret_cast<GemmWsSmem*>(gemm_ws_smem_raw + ((1024u - (smem_u32(gemm_ws_smem_raw) & 1023u)) & 1023u));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int nb = blockIdx.x % n_blocks, grp = blockIdx.x / n_blocks;
  const int n0 = nb * WBN;
  const int m_tiles = (M + WBM - 1) / WBM;
  const int my_tiles = grp < m_tiles ? (m_tiles - grp + groups - 1) / groups : 0;
  const int nkb = (nk16 + 3) >> 2;
  const int n_st = nkb <= 4 ? WST_MAX : WST;
  auto stage = [&](int s) -> __half* { return s < WST ? sm.a[s] : sm.w[WKB_MAX - 1] + (s - WST) * (WBM * 64); };
  for (int i = threadIdx.x; i < WBN; i += WTHREADS) sm.bias[i] = bias ? __ldg(bias + n0 + i) : 0.f;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    mbar_init(&sm.w_full, 1);
    for (int s = 0; s < WST_MAX; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int b = 0; b < WACC; ++b) { mbar_init(&sm.acc_full[b], 1); mbar_init(&sm.acc_empty[b], 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&sm.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&sm.w_full, (uint32_t)(nkb * sizeof(sm.w[0])));
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sm.w[kb], &tmW, &sm.w_full, kb * 64, n0);
      // A tiles come from HBM for the first of the n_blocks CTAs that touches them, while HBM is busy absorbing C
      // (ncu r2h: tensor pipe 50 % active, the MMA thread waiting for A): the CTA whose column block matches the
      // tile index pulls the tile WPF iterations ahead into L2
      constexpr int WPF = 3;
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int m0 = (grp + i * groups) * WBM;
        if (i + WPF < my_tiles && (i + WPF) % n_blocks == nb) {
          const int mp = (grp + (i + WPF) * groups) * WBM;
          for (int kb = 0; kb < nkb; ++kb) tma_prefetch_l2_2d(&tmA, kb * 64, mp);
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % n_st;
          mbar_wait(&sm.empty[s], ((it / n_st) & 1) ^ 1);
          mbar_expect_tx(&sm.full[s], (uint32_t)sizeof(sm.a[0]));
          tma_load_2d(stage(s), &tmA, &sm.full[s], kb * 64, m0);
        }
      }
    }
  } else if (warp == 1) {
    {   // the whole warp, converged: one elected lane issues (tc.cuh, umma_*_elect)
      constexpr uint32_t idesc = umma_idesc(FMT_F16, WBM, WBN);
      mbar_wait(&sm.w_full, 0);
      tc_fence_after();
      // debug (AVHOT_GEMM_TRACE=<file>): CTA 0's MMA warp sums, over tiles [64, 192), the cycles it waited for a free
      // accumulator and for A tiles, and the total
      const bool probe = trace && blockIdx.x == 0;
      long long w_acc = 0, w_full = 0, t_begin = 0;
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int b = i % WACC;
        const bool tr = probe && i >= 600 && i < 728;
        if (probe && i == 600) t_begin = clock64();
        long long c0 = tr ? clock64() : 0;
        mbar_wait(&sm.acc_empty[b], ((uint32_t)(i / WACC) & 1) ^ 1);
        tc_fence_after();
        if (tr) w_acc += clock64() - c0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % n_st;
          c0 = tr ? clock64() : 0;
          mbar_wait(&sm.full[s], (it / n_st) & 1);
          tc_fence_after();
          if (tr) w_full += clock64() - c0;
          const uint64_t adesc = umma_desc_sw128(smem_u32(stage(s)));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sm.w[kb]));
          const int nk = min(4, nk16 - 4 * kb);
          for (int k = 0; k < nk; ++k)
            umma_f16_elect(tmem + (uint32_t)(b * WBN), adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_elect(&sm.empty[s]);
        }
        umma_commit_elect(&sm.acc_full[b]);
        if (probe && i == 727 && lane == 0) { trace[0] = w_acc; trace[1] = w_full; trace[2] = clock64() - t_begin; }
      }
    }
  } else {
    const int wg = (warp - 2) >> 2, quarter = warp & 3;
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
    long long e_wait = 0, e_work = 0;
    const bool eprobe = trace && blockIdx.x == 0 && warp == 2;
    for (int i = wg; i < my_tiles; i += 2) {
      const int b = i % WACC;
      const int row = (grp + i * groups) * WBM + quarter * 32 + lane;
      const bool tr = eprobe && i >= 600 && i < 728;
      const long long c0 = tr ? clock64() : 0;
      mbar_wait(&sm.acc_full[b], (uint32_t)(i / WACC) & 1);
      tc_fence_after();
      const long long c1 = tr ? clock64() : 0;
      e_wait += c1 - c0;
      // Output layout [M / 128][N][128] ("tile-transposed"): element (row, col) of M-tile mt lives at
      // ((mt * N + col) * 128 + row % 128), so for every column the 32 lanes of a warp write 128 contiguous bytes
      // (1 wavefront per store instead of 32 for row-major rows that are 4 KB apart), and the consumer
      // (lstm_tc.cu, thread = row) reads it back the same way.  Pure TMEM -> global copy, no arithmetic.
      if (!OUT_F16) {
        float* cbase = C + ((size_t)(row / WBM) * ldc + n0) * WBM + (quarter * 32 + lane);
#pragma unroll
        for (int c = 0; c < WBN; c += 64) {
          uint32_t r0[32], r1[32];
          tmem_ld32(tlane + (uint32_t)(b * WBN + c), r0);
          tmem_ld32(tlane + (uint32_t)(b * WBN + c + 32), r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) cbase[(size_t)(c + j) * WBM] = __uint_as_float(r0[j]) + sm.bias[c + j];
#pragma unroll
          for (int j = 0; j < 32; ++j) cbase[(size_t)(c + 32 + j) * WBM] = __uint_as_float(r1[j]) + sm.bias[c + 32 + j];
        }
      } else {
        // fp16 output: groups of 8 consecutive columns of a row packed into 16 bytes ->
        // [M / 128][N / 8][128 rows][8 halves]: a warp stores 512 contiguous bytes per instruction and the consumer
        // (lstm_tc.cu, thread = row) reads the 8 units of a gate with one 128-bit load
        uint4* cbase = reinterpret_cast<uint4*>(C) + ((size_t)(row / WBM) * (ldc / 8) + n0 / 8) * WBM + (quarter * 32 + lane);
#pragma unroll
        for (int c = 0; c < WBN; c += 64) {
          uint32_t r0[32], r1[32];
          tmem_ld32(tlane + (uint32_t)(b * WBN + c), r0);
          tmem_ld32(tlane + (uint32_t)(b * WBN + c + 32), r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 bb = *reinterpret_cast<const float2*>(&sm.bias[c + j + 2 * e]);
              const __half2 v = __floats2half2_rn(__uint_as_float(r0[j + 2 * e]) + bb.x, __uint_as_float(r0[j + 2 * e + 1]) + bb.y);
              pk[e] = *reinterpret_cast<const uint32_t*>(&v);
            }
            cbase[(size_t)((c + j) / 8) * WBM] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 bb = *reinterpret_cast<const float2*>(&sm.bias[c + 32 + j + 2 * e]);
              const __half2 v = __floats2half2_rn(__uint_as_float(r1[j + 2 * e]) + bb.x, __uint_as_float(r1[j + 2 * e + 1]) + bb.y);
              pk[e] = *reinterpret_cast<const uint32_t*>(&v);
            }
            cbase[(size_t)((c + 32 + j) / 8) * WBM] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.acc_empty[b]);      // one arrival per epilogue warp
      if (tr) e_work += clock64() - c1;
      if (eprobe && i == 726 && lane == 0) { trace[3] = e_wait; trace[4] = e_work; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}


// ---- CTA-pair form (cta_group::2) -------------------------------------------------------------------------------------
// The AVHOT_GEMM_TRACE wait counters showed what bounds the single-CTA kernel at K = 320: its MMA warp waits 2 000 of a
// tile's 4 500 cycles for A tiles - the 160 KB weight block leaves room for a ring of only four 16 KB stages, fewer bytes
// in flight than the L2 latency needs.  A CTA pair shares the weight block: each CTA keeps 128 of the 256 columns (80 KB)
// and its own 128-row A tile, the leader issues tcgen05.mma.cta_group::2 (M 256 = both tiles, N 256 = both halves of W),
// and the freed 80 KB become an eight-stage A ring in each CTA.  TMA completions of both CTAs land on the leader's `full`
// barriers, MMA completions are multicast to both CTAs' `empty` / `acc_full` barriers, both CTAs' epilogue warps arrive on
// the leader's `acc_empty` barriers.
constexpr int PST = 8, WHALF = WBN / 2;

struct GemmWsPairSmem {
  __half w[WKB_MAX][WHALF * 64];   // this CTA's 128 columns of the pair's weight block, 16 KB per 64-wide k-block
  __half a[PST][WBM * 64];         // A ring
  float bias[WBN];
  uint64_t w_full, full[PST], empty[PST], acc_full[WACC], acc_empty[WACC];
  uint32_t tmem_base;
};

template <bool OUT_F16>
__global__ void __launch_bounds__(WTHREADS, 1)
gemm_f16_ws_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWh,
                        const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int nk16, int n_blocks,
                        int groups, long long* __restrict__ trace) {
  extern __shared__ unsigned char gemm_wsp_smem_raw[];
  GemmWsPairSmem& sm = *reinterpret_cast<GemmWsPairSmem*>(gemm_wsp_smem_raw + ((1024u - (smem_u32(gemm_wsp_smem_raw) & 1023u)) & 1023u));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int pair = (int)(blockIdx.x >> 1);
  const int nb = pair % n_blocks, grp = pair / n_blocks;
  const int n0 = nb * WBN;
  const int m_tiles = M / WBM, m_pairs = (m_tiles + 1) / 2;
  const int my_pairs = grp < m_pairs ? (m_pairs - grp + groups - 1) / groups : 0;
  const int nkb = (nk16 + 3) >> 2;
  for (int i = threadIdx.x; i < WBN; i += WTHREADS) sm.bias[i] = bias ? __ldg(bias + n0 + i) : 0.f;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmWh);
    mbar_init(&sm.w_full, 1);
    for (int s = 0; s < PST; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int b = 0; b < WACC; ++b) { mbar_init(&sm.acc_full[b], 1); mbar_init(&sm.acc_empty[b], 8); }   // 4 warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(&sm.tmem_base, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // both CTAs' barriers are initialised before any remote arrival
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t wfull0 = cluster_map_shared(&sm.w_full, 0), full0 = cluster_map_shared(&sm.full[0], 0);
      if (leader) mbar_expect_tx(&sm.w_full, (uint32_t)(2 * nkb * sizeof(sm.w[0])));
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d_pair(sm.w[kb], &tmWh, wfull0, kb * 64, n0 + (int)crank * WHALF);
      // no L2 prefetch here (the single-CTA kernel needs one three tiles ahead): with the eight-stage ring it only made
      // every A tile come from DRAM twice (ncu r5c: 7.5 GB read per launch; 3.65 GB = the operand once without it, and
      // the launch 7 % faster)
      uint32_t it = 0;
      for (int i = 0; i < my_pairs; ++i) {
        const int m0 = (2 * (grp + i * groups) + (int)crank) * WBM;       // rows >= M read as zeros
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % PST;
          mbar_wait(&sm.empty[s], ((it / PST) & 1) ^ 1);
          if (leader) mbar_expect_tx(&sm.full[s], (uint32_t)(2 * sizeof(sm.a[0])));   // both CTAs' tiles
          tma_load_2d_pair(sm.a[s], &tmA, full0 + (uint32_t)s * 8u, kb * 64, m0);
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {                      // the whole warp, converged; one elected lane issues
      constexpr uint32_t idesc = umma_idesc(FMT_F16, 2 * WBM, WBN);
      mbar_wait(&sm.w_full, 0);
      tc_fence_after();
      const bool probe = trace && blockIdx.x == 0;     // AVHOT_GEMM_TRACE: as in the single-CTA kernel
      long long w_acc = 0, w_full = 0, t_begin = 0;
      uint32_t it = 0;
      for (int i = 0; i < my_pairs; ++i) {
        const int b = i % WACC;
        const bool tr = probe && i >= 600 && i < 728;
        if (probe && i == 600) t_begin = clock64();
        long long c0 = tr ? clock64() : 0;
        mbar_wait_cluster(&sm.acc_empty[b], ((uint32_t)(i / WACC) & 1) ^ 1);
        tc_fence_after();
        if (tr) w_acc += clock64() - c0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % PST;
          c0 = tr ? clock64() : 0;
          mbar_wait(&sm.full[s], (it / PST) & 1);
          tc_fence_after();
          if (tr) w_full += clock64() - c0;
          const uint64_t adesc = umma_desc_sw128(smem_u32(sm.a[s]));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sm.w[kb]));
          const int nk = min(4, nk16 - 4 * kb);
          for (int k = 0; k < nk; ++k)
            umma_f16_pair_elect(tmem + (uint32_t)(b * WBN), adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair_mc_elect(&sm.empty[s], 3);
        }
        umma_commit_pair_mc_elect(&sm.acc_full[b], 3);
        if (probe && i == 727 && lane == 0) { trace[0] = w_acc; trace[1] = w_full; trace[2] = clock64() - t_begin; }
      }
    }
  } else {
    const int wg = (warp - 2) >> 2, quarter = warp & 3;
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t acc_empty0 = cluster_map_shared(&sm.acc_empty[0], 0);
    long long e_wait = 0, e_work = 0;
    const bool eprobe = trace && blockIdx.x == 0 && warp == 2;
    for (int i = wg; i < my_pairs; i += 2) {
      const int b = i % WACC;
      const int tile = 2 * (grp + i * groups) + (int)crank;
      const int row = tile * WBM + quarter * 32 + lane;
      const bool tr = eprobe && i >= 600 && i < 728;
      const long long c0 = tr ? clock64() : 0;
      mbar_wait(&sm.acc_full[b], (uint32_t)(i / WACC) & 1);
      tc_fence_after();
      const long long c1 = tr ? clock64() : 0;
      e_wait += c1 - c0;
      if (tile < m_tiles) {
        if (!OUT_F16) {
          float* cbase = C + ((size_t)tile * ldc + n0) * WBM + (quarter * 32 + lane);
#pragma unroll
          for (int c = 0; c < WBN; c += 64) {
            uint32_t r0[32], r1[32];
            tmem_ld32(tlane + (uint32_t)(b * WBN + c), r0);
            tmem_ld32(tlane + (uint32_t)(b * WBN + c + 32), r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) cbase[(size_t)(c + j) * WBM] = __uint_as_float(r0[j]) + sm.bias[c + j];
#pragma unroll
            for (int j = 0; j < 32; ++j) cbase[(size_t)(c + 32 + j) * WBM] = __uint_as_float(r1[j]) + sm.bias[c + 32 + j];
          }
        } else {
          uint4* cbase = reinterpret_cast<uint4*>(C) + ((size_t)tile * (ldc / 8) + n0 / 8) * WBM + (quarter * 32 + lane);
#pragma unroll
          for (int c = 0; c < WBN; c += 64) {
            uint32_t r0[32], r1[32];
            tmem_ld32(tlane + (uint32_t)(b * WBN + c), r0);
            tmem_ld32(tlane + (uint32_t)(b * WBN + c + 32), r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 bb = *reinterpret_cast<const float2*>(&sm.bias[c + j + 2 * e]);
                const __half2 v = __floats2half2_rn(__uint_as_float(r0[j + 2 * e]) + bb.x, __uint_as_float(r0[j + 2 * e + 1]) + bb.y);
                pk[e] = *reinterpret_cast<const uint32_t*>(&v);
              }
              cbase[(size_t)((c + j) / 8) * WBM] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 bb = *reinterpret_cast<const float2*>(&sm.bias[c + 32 + j + 2 * e]);
                const __half2 v = __floats2half2_rn(__uint_as_float(r1[j + 2 * e]) + bb.x, __uint_as_float(r1[j + 2 * e + 1]) + bb.y);
                pk[e] = *reinterpret_cast<const uint32_t*>(&v);
              }
              cbase[(size_t)((c + 32 + j) / 8) * WBM] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      }
      (void)row;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty0 + (uint32_t)b * 8u);   // one arrival per epilogue warp, on the leader
      if (tr) e_work += clock64() - c1;
      if (eprobe && i == 726 && lane == 0) { trace[3] = e_wait; trace[4] = e_work; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // no CTA leaves while its peer can still signal its barriers
  if (warp == 1) tmem_dealloc_pair(tmem, 512);
}

}  // namespace tc

// a, w fp16 row-major (K contiguous); M % 128 == 0 (caller pads); N % 256 == 0; K <= 320;
// c in the tile-transposed layout: fp32 [M / 128][N = ldc][128], or fp16 column octets [M / 128][N / 8][128][8]
int gemm_f16_ws(const void* a, int lda, const void* w, int ldw, const float* bias, void* c_out, bool out_f16, int ldc,
                int M, int N, int K, cudaStream_t s, const char* prof_name) {
  float* c = reinterpret_cast<float*>(c_out);
  using namespace tc;
  AVHOT_REQUIRE(a && w && c && M >= 0 && N > 0 && K > 0, AVHOT_ERR_INVALID, "gemm_ws: bad arguments");
  AVHOT_REQUIRE(M % WBM == 0 && N % WBN == 0 && ldc == N && K <= 64 * WKB_MAX && (lda & 7) == 0 && (ldw & 7) == 0,
                AVHOT_ERR_UNSUPPORTED, "gemm_ws: needs M %% 128 == 0, N %% 256 == 0, ldc == N, K <= 320, lda/ldw %% 8 == 0 "
                "(M=%d N=%d K=%d)", M, N, K);
  if (M == 0) return AVHOT_OK;
  CUtensorMap tmA, tmW;
  int rc = make_tensor_map_2d(&tmA, a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda, WBM, 64);
  if (rc != AVHOT_OK) return rc;
  rc = make_tensor_map_2d(&tmW, w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, WBN, 64);
  if (rc != AVHOT_OK) return rc;
  const int n_blocks = N / WBN;
  const int m_tiles = ceil_div(M, WBM);
  static const char* trace_path = getenv("AVHOT_GEMM_TRACE");
  long long* d_trace = nullptr;
  if (trace_path && m_tiles >= 200 * num_sms() / n_blocks) {
    AVHOT_CUDA(cudaMalloc(&d_trace, 8 * sizeof(long long)));
    AVHOT_CUDA(cudaMemsetAsync(d_trace, 0, 8 * sizeof(long long), s));
  }
  auto dump_trace = [&](const char* form) -> int {
    if (!d_trace) return AVHOT_OK;
    long long h[8];
    AVHOT_CUDA(cudaMemcpyAsync(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost, s));
    AVHOT_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_trace);
    if (FILE* f = fopen(trace_path, "a")) {
      fprintf(f, "%s K=%d per tile (128 tiles): mma total %lld wait_acc_empty %lld wait_A %lld | epilogue (64 tiles of one warpgroup): "
                 "wait_acc_full %lld work %lld\n", form, K, h[2] / 128, h[0] / 128, h[1] / 128, h[3] / 64, h[4] / 64);
      fclose(f);
    }
    return AVHOT_OK;
  };
  // CTA pairs (cta_group::2) when every pair of the grid can be resident at once; AVHOT_GEMM_PAIR=0 keeps the single-CTA form
  static const bool want_pair = [] { const char* e = getenv("AVHOT_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  const int pair_groups = (num_sms() / 2) / n_blocks;
  if (want_pair && pair_groups >= 1 && m_tiles >= 4 * pair_groups) {
    CUtensorMap tmWh;
    rc = make_tensor_map_2d(&tmWh, w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, WHALF, 64);
    if (rc != AVHOT_OK) return rc;
    const int psmem = (int)sizeof(GemmWsPairSmem) + 1024;
    auto kern = out_f16 ? gemm_f16_ws_pair_kernel<true> : gemm_f16_ws_pair_kernel<false>;
    AVHOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem));
    AVHOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 0));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * n_blocks * pair_groups));
    cfg.blockDim = dim3(WTHREADS);
    cfg.dynamicSmemBytes = (size_t)psmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters >= n_blocks * pair_groups) {
      const int nk16p = ceil_div(K, 16);
      {
        ProfScope prof(prof_name, s);
        AVHOT_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmWh, bias, c, ldc, M, nk16p, n_blocks, pair_groups, d_trace));
        count_launch();
      }
      return dump_trace("pair");
    }
    (void)cudaGetLastError();
  }
  int groups = num_sms() / n_blocks;
  if (groups < 1) groups = 1;
  if (groups > m_tiles) groups = m_tiles;
  const int nk16 = ceil_div(K, 16);
  const int smem = (int)sizeof(GemmWsSmem) + 1024;
  {
    ProfScope prof(prof_name, s);
    if (out_f16) {
      AVHOT_CUDA(cudaFuncSetAttribute(gemm_f16_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      gemm_f16_ws_kernel<true><<<n_blocks * groups, WTHREADS, smem, s>>>(tmA, tmW, bias, c, ldc, M, nk16, n_blocks, groups, d_trace);
    } else {
      AVHOT_CUDA(cudaFuncSetAttribute(gemm_f16_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      gemm_f16_ws_kernel<false><<<n_blocks * groups, WTHREADS, smem, s>>>(tmA, tmW, bias, c, ldc, M, nk16, n_blocks, groups, d_trace);
    }
    AVHOT_LAUNCH_CHECK();
  }
  return dump_trace("single");
}

}  // namespace avhot
