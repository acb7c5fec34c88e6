// tcgen05 GEMM:  C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]),  A and W fp32 in HBM, multiplied as TF32 on the
// 5th-generation tensor cores with fp32 accumulation in TMEM  (AVHOT_PREC_TC path of the time-batched
// contractions: video Linear(12544 -> 256), LSTM input projections).
//
// fp32 operands are fed to the tensor core as they are (kind::tf32 reads fp32 words from shared memory and
// uses the top 19 bits), so A is read from HBM exactly once and never re-written in another type.
//
// One CTA = one 128 x BN output tile, 192 threads, warp-specialised:
//   warp 0   TMA producer: 128 x 32 (A) and BN x 32 (W) fp32 boxes (128-byte rows, 128-byte swizzle) into a
//            ring of stages, each completing an mbarrier with its byte count;
//   warp 1   allocates BN TMEM columns, then one lane issues tcgen05.mma (M 128, N BN, K 8) x 4 per stage and
//            tcgen05.commit's the stage back to the producer; a final commit signals the accumulator;
//   warps 2-5  epilogue: tcgen05.ld of the warp's 32-lane TMEM quarter (thread = output row), bias +
//            activation, 128-bit stores.
// Two CTAs fit per SM (<= 113 KB shared memory and <= 256 TMEM columns each), so one tile's epilogue overlaps
// the other's main loop without a persistent scheduler.
#include <mutex>

#include "common.cuh"
#include "tc.cuh"

namespace avhot {
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static void resolve_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
}

int make_tensor_map_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int elem_bytes, uint64_t rows,
                       uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  resolve_encode();
  AVHOT_REQUIRE(g_encode, AVHOT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  AVHOT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * elem_bytes) % 16 == 0, AVHOT_ERR_INVALID,
                "tensor map: base must be 16-byte aligned and the row pitch a multiple of 16 bytes");
  AVHOT_REQUIRE(box_cols * elem_bytes == 128 && box_rows <= 256, AVHOT_ERR_INVALID, "tensor map: bad box");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode(out, dtype, 2, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AVHOT_REQUIRE(r == CUDA_SUCCESS, AVHOT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return AVHOT_OK;
}

int make_tensor_map_3d(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int elem_bytes,
                       const uint64_t (&dims)[3], const uint64_t (&strides_bytes)[2], const uint32_t (&box)[3]) {
  resolve_encode();
  AVHOT_REQUIRE(g_encode, AVHOT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  AVHOT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && strides_bytes[0] % 16 == 0 && strides_bytes[1] % 16 == 0,
                AVHOT_ERR_INVALID, "tensor map: base and strides must be multiples of 16 bytes");
  AVHOT_REQUIRE(box[0] * (uint32_t)elem_bytes == 128 && box[1] <= 256 && box[2] <= 256, AVHOT_ERR_INVALID,
                "tensor map: bad box");
  const cuuint64_t d[3] = {dims[0], dims[1], dims[2]};
  const cuuint64_t st[2] = {strides_bytes[0], strides_bytes[1]};
  const cuuint32_t bx[3] = {box[0], box[1], box[2]};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult rc = g_encode(out, dtype, 3, const_cast<void*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AVHOT_REQUIRE(rc == CUDA_SUCCESS, AVHOT_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)rc);
  return AVHOT_OK;
}

constexpr int TBM = 128, TBK = 32, TSTAGES = 2, TTHREADS = 192;

template <int BN>
struct GemmTcSmem {
  float a[TSTAGES][TBM * TBK];   // 16 KB per stage, 1024-byte aligned
  float w[TSTAGES][BN * TBK];    // BN * 128 bytes per stage
  uint64_t full[TSTAGES], empty[TSTAGES], acc_full;
  uint32_t tmem_base;
};

template <int BN>
__global__ void __launch_bounds__(TTHREADS)
gemm_tf32_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N, int K, int act) {
  extern __shared__ unsigned char gemm_tc_smem_raw[];
  GemmTcSmem<BN>& sm = *reinterpret_cast<GemmTcSmem<BN>*>(
      (reinterpret_cast<uintptr_t>(gemm_tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * BN;
  const int nkb = (K + TBK - 1) / TBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < TSTAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&sm.tmem_base, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TSTAGES;
        const uint32_t phase = (kb / TSTAGES) & 1;
        mbar_wait(&sm.empty[s], phase ^ 1);            // a fresh barrier passes the parity-1 wait
        mbar_expect_tx(&sm.full[s], (uint32_t)((TBM + BN) * TBK * sizeof(float)));
        tma_load_2d(sm.a[s], &tmA, &sm.full[s], kb * TBK, m0);
        tma_load_2d(sm.w[s], &tmW, &sm.full[s], kb * TBK, n0);
      }
    }
  } else if (warp == 1) {
    {   // the whole warp, converged: one elected lane issues (tc.cuh, umma_*_elect)
      constexpr uint32_t idesc = umma_idesc(FMT_TF32, TBM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TSTAGES;
        const uint32_t phase = (kb / TSTAGES) & 1;
        mbar_wait(&sm.full[s], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(sm.a[s]));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sm.w[s]));
#pragma unroll
        for (int k = 0; k < TBK / 8; ++k)               // 8 fp32 = 32 bytes per MMA along K: +2 in the >>4 address
          umma_tf32_elect(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        umma_commit_elect(&sm.empty[s]);                      // stage free once these MMAs have read it
      }
      umma_commit_elect(&sm.acc_full);
    }
  } else {
    const int quarter = warp & 3;                       // the TMEM lanes this warp may access
    const int row = m0 + quarter * 32 + lane;
    mbar_wait(&sm.acc_full, 0);
    tc_fence_after();
    float* crow = C + (size_t)row * ldc + n0;
    const bool vec_ok = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, r);
      tmem_ld_wait();
      if (row < M) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + c + j;
          float v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float x = __uint_as_float(r[j + u]);
            if (bias && n + u < N) x += __ldg(bias + n + u);
            if (act == ACT_RELU) x = fmaxf(x, 0.f);
            v[u] = x;
          }
          if (vec_ok && n + 3 < N) {
            *reinterpret_cast<float4*>(crow + c + j) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (n + u < N) crow[c + j + u] = v[u];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, BN);
}

}  // namespace tc

// host entry: same contract as gemm_f32 (fp32 A only)
int gemm_tf32_tc(const float* a, int lda, const float* w, int ldw, const float* bias, float* c, int ldc, int M, int N,
                 int K, int act, cudaStream_t s, const char* prof_name) {
  using namespace tc;
  AVHOT_REQUIRE(a && w && c && M >= 0 && N > 0 && K > 0, AVHOT_ERR_INVALID, "gemm_tc: bad arguments");
  AVHOT_REQUIRE((lda & 3) == 0 && (ldw & 3) == 0, AVHOT_ERR_INVALID, "gemm_tc: lda, ldw must be multiples of 4");
  if (M == 0) return AVHOT_OK;
  const int BN = N > 128 ? 256 : 128;
  CUtensorMap tmA, tmW;
  int rc = make_tensor_map_2d(&tmA, a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)M, (uint64_t)K, (uint64_t)lda, TBM, TBK);
  if (rc != AVHOT_OK) return rc;
  rc = make_tensor_map_2d(&tmW, w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BN, TBK);
  if (rc != AVHOT_OK) return rc;
  dim3 grid(ceil_div(N, BN), ceil_div(M, TBM));
  AVHOT_REQUIRE(grid.y <= 65535, AVHOT_ERR_INVALID, "gemm_tc: M too large for one launch");
  ProfScope prof(prof_name, s);
  if (BN == 256) {
    const int smem = (int)sizeof(GemmTcSmem<256>) + 1024;
    AVHOT_CUDA(cudaFuncSetAttribute(gemm_tf32_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gemm_tf32_tc_kernel<256><<<grid, TTHREADS, smem, s>>>(tmA, tmW, bias, c, ldc, M, N, K, act);
  } else {
    const int smem = (int)sizeof(GemmTcSmem<128>) + 1024;
    AVHOT_CUDA(cudaFuncSetAttribute(gemm_tf32_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gemm_tf32_tc_kernel<128><<<grid, TTHREADS, smem, s>>>(tmA, tmW, bias, c, ldc, M, N, K, act);
  }
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

}  // namespace avhot
