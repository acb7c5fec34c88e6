// Weight-stationary tcgen05 GEMM for the LSTM input projections (AVHOT_PREC_TC path):
//     C[M, N] fp32 = A[M, K] (fp16) * W[N, K]^T (fp16),   M = B*T ~ 10^6 rows, N = 1024, K <= 320.
// The bias rides in the GEMM itself: A carries two columns of ones against (bias_hi, bias_lo) columns of W, so the
// epilogue - the bottleneck of this output-bound op - is a pure copy.
// The op is bound by writing C (4 KB per row against <= 640 B of A), so the design minimises everything else:
// a persistent CTA owns one 128-column block of W for its whole life (<= 80 KB of fp16 in shared memory, loaded
// once by TMA) and walks down M; A tiles (128 rows x 64 k, 128-byte swizzle) stream through an 8-stage TMA ring;
// accumulators rotate through four 128-column TMEM buffers so the epilogue of tile i overlaps the MMAs of tiles
// i+1..i+3.  The 8 CTAs that share an M tile (one per column block) walk the same tile sequence, so A is read
// from HBM once and served to the other seven from L2.
//   warp 0     TMA producer (W block once, then the A ring)
//   warp 1     TMEM allocation + one lane issuing tcgen05.mma kind::f16 (M 128, N 128, K 16)
//   warps 2-9  two epilogue warpgroups (even / odd tiles): tcgen05.ld (thread = row) -> 256-bit global stores
// Layer 0's A is the fp16 3-way split [hi | lo | hi] of the features against W = [hi | hi | lo] (see lstm.cu), which
// makes this single fp16 pass fp32-class; layer 1's A is the fp16 hidden sequence written by the recurrent kernel.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc.cuh"

namespace avhot {
namespace tc {

constexpr int WBM = 128, WBN = 128, WKB_MAX = 5, WST = 8, WACC = 4, WTHREADS = 320;

struct GemmWsSmem {
  __half w[WKB_MAX][WBN * 64];   // the CTA's weight block, one 16 KB tile per 64-wide k-block
  __half a[WST][WBM * 64];       // A ring
  uint64_t w_full, full[WST], empty[WST], acc_full[WACC], acc_empty[WACC];
  uint32_t tmem_base;
};

template <bool OUT_F16>
__global__ void __launch_bounds__(WTHREADS, 1)
gemm_f16_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                   float* __restrict__ C, int ldc, int M, int nkb, int n_blocks,
                   int groups) {
  extern __shared__ unsigned char gemm_ws_smem_raw[];
  GemmWsSmem& sm = *reinterpret_cast<GemmWsSmem*>((reinterpret_cast<uintptr_t>(gemm_ws_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = blockIdx.x % n_blocks, grp = blockIdx.x / n_blocks;
  const int n0 = nb * WBN;
  const int m_tiles = (M + WBM - 1) / WBM;
  const int my_tiles = grp < m_tiles ? (m_tiles - grp + groups - 1) / groups : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    mbar_init(&sm.w_full, 1);
    for (int s = 0; s < WST; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int b = 0; b < WACC; ++b) { mbar_init(&sm.acc_full[b], 1); mbar_init(&sm.acc_empty[b], 128); }
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
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int m0 = (grp + i * groups) * WBM;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % WST;
          mbar_wait(&sm.empty[s], ((it / WST) & 1) ^ 1);
          mbar_expect_tx(&sm.full[s], (uint32_t)sizeof(sm.a[0]));
          tma_load_2d(sm.a[s], &tmA, &sm.full[s], kb * 64, m0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(FMT_F16, WBM, WBN);
      mbar_wait(&sm.w_full, 0);
      tc_fence_after();
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int b = i % WACC;
        mbar_wait(&sm.acc_empty[b], ((uint32_t)(i / WACC) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % WST;
          mbar_wait(&sm.full[s], (it / WST) & 1);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sm.a[s]));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sm.w[kb]));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem + (uint32_t)(b * WBN), adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&sm.empty[s]);
        }
        umma_commit(&sm.acc_full[b]);
      }
    }
  } else {
    const int wg = (warp - 2) >> 2, quarter = warp & 3;
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int i = wg; i < my_tiles; i += 2) {
      const int b = i % WACC;
      const int row = (grp + i * groups) * WBM + quarter * 32 + lane;
      mbar_wait(&sm.acc_full[b], (uint32_t)(i / WACC) & 1);
      tc_fence_after();
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
          for (int j = 0; j < 32; ++j) cbase[(size_t)(c + j) * WBM] = __uint_as_float(r0[j]);
#pragma unroll
          for (int j = 0; j < 32; ++j) cbase[(size_t)(c + 32 + j) * WBM] = __uint_as_float(r1[j]);
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
              const __half2 v = __floats2half2_rn(__uint_as_float(r0[j + 2 * e]), __uint_as_float(r0[j + 2 * e + 1]));
              pk[e] = *reinterpret_cast<const uint32_t*>(&v);
            }
            cbase[(size_t)((c + j) / 8) * WBM] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __half2 v = __floats2half2_rn(__uint_as_float(r1[j + 2 * e]), __uint_as_float(r1[j + 2 * e + 1]));
              pk[e] = *reinterpret_cast<const uint32_t*>(&v);
            }
            cbase[(size_t)((c + 32 + j) / 8) * WBM] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&sm.acc_empty[b]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace tc

// a, w fp16 row-major (K contiguous); M % 128 == 0 (caller pads); N % 128 == 0; K <= 320;
// c in the tile-transposed layout: fp32 [M / 128][N = ldc][128], or fp16 column octets [M / 128][N / 8][128][8]
int gemm_f16_ws(const void* a, int lda, const void* w, int ldw, void* c_out, bool out_f16, int ldc, int M, int N,
                int K, cudaStream_t s, const char* prof_name) {
  float* c = reinterpret_cast<float*>(c_out);
  using namespace tc;
  AVHOT_REQUIRE(a && w && c && M >= 0 && N > 0 && K > 0, AVHOT_ERR_INVALID, "gemm_ws: bad arguments");
  AVHOT_REQUIRE(M % WBM == 0 && N % WBN == 0 && ldc == N && K <= 64 * WKB_MAX && (lda & 7) == 0 && (ldw & 7) == 0,
                AVHOT_ERR_UNSUPPORTED, "gemm_ws: needs M, N %% 128 == 0, ldc == N, K <= 320, lda/ldw %% 8 == 0 "
                "(M=%d N=%d K=%d)", M, N, K);
  if (M == 0) return AVHOT_OK;
  CUtensorMap tmA, tmW;
  int rc = make_tensor_map_2d(&tmA, a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda, WBM, 64);
  if (rc != AVHOT_OK) return rc;
  rc = make_tensor_map_2d(&tmW, w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, WBN, 64);
  if (rc != AVHOT_OK) return rc;
  const int n_blocks = N / WBN;
  const int m_tiles = ceil_div(M, WBM);
  int groups = num_sms() / n_blocks;
  if (groups < 1) groups = 1;
  if (groups > m_tiles) groups = m_tiles;
  const int nkb = ceil_div(K, 64);
  const int smem = (int)sizeof(GemmWsSmem) + 1024;
  ProfScope prof(prof_name, s);
  if (out_f16) {
    AVHOT_CUDA(cudaFuncSetAttribute(gemm_f16_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gemm_f16_ws_kernel<true><<<n_blocks * groups, WTHREADS, smem, s>>>(tmA, tmW, c, ldc, M, nkb, n_blocks, groups);
  } else {
    AVHOT_CUDA(cudaFuncSetAttribute(gemm_f16_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gemm_f16_ws_kernel<false><<<n_blocks * groups, WTHREADS, smem, s>>>(tmA, tmW, c, ldc, M, nkb, n_blocks, groups);
  }
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

}  // namespace avhot
