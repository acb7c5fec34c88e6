// uint8 frames straight onto the tensor cores (AVHOT_PREC_TC path of FrameEncoder's Linear + ReLU when the caller hands
// over the decoded pixels, src/dataprocessing.py:261-265 divides them by 255 on the host):
//     C[M, 256] = relu( (A_u8[M, K] x Wq[256, K]^T) / s + bias ),   Wq = fp16(W s / 255),  s = a power of two
// A pixel k is exact in fp16, so the product needs one fp16 pass (the weights carry 11 significant bits - one more than
// the TF32 path gives both operands) and A costs a quarter of the fp32 frames' HBM bytes: 200 704 B per clip instead of
// 802 816 B, and no fp32 image of the pixels is ever written.
//
// Persistent CTA (one per SM, grid-stride over 128-row tiles), 12 warps:
//   warp 0      TMA producer: raw u8 boxes (128 rows x 128 B = two k-blocks, 128-byte swizzle) through a 4-stage ring,
//               fp16 weight tiles (256 rows x 64 k) through a 3-stage ring (the 6.4 MB weight matrix stays in L2);
//   warps 8-11  converters, thread = row: 16 pixels per 128-bit shared load, bytes -> fp16 by PRMT (0x6400 | k = 1024 + k)
//               and one HSUB2, written as the A operand's K-major swizzled k-block (3-stage ring);
//   warp 1      one elected lane issues tcgen05.mma kind::f16 (M 128, N 256, K 16) x 4 per k-block into one of two
//               256-column TMEM accumulators and commits the converted / weight stages back;
//   warps 4-7   epilogue of the previous tile while the next one accumulates: tcgen05.ld (thread = row), x 1 / s, bias,
//               ReLU, 128-bit stores.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc.cuh"

namespace avhot {
namespace tc {

constexpr int UBM = 128, UBN = 256, URAW = 4, UCV = 3, UWS = 3, UTHREADS = 384;

struct GemmU8Smem {
  uint8_t raw[URAW][UBM * 128];     // 128 rows x 128 pixels (two k-blocks of 64), as TMA wrote them
  __half cv[UCV][UBM * 64];         // converted A k-block
  __half w[UWS][UBN * 64];          // weight k-block
  float bias[UBN];
  uint64_t raw_full[URAW], raw_empty[URAW], cv_full[UCV], cv_empty[UCV], w_full[UWS], w_empty[UWS], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(GemmU8Smem) + 1024 <= 227 * 1024, "one CTA per SM");

__global__ void __launch_bounds__(UTHREADS, 1)
gemm_u8_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                   const float* __restrict__ bias, const float* __restrict__ inv_scale, float* __restrict__ C, int ldc, int M,
                   int K, int act) {
  extern __shared__ unsigned char gemm_u8_smem_raw[];
  GemmU8Smem& sm = *reinterpret_cast<GemmU8Smem*>(gemm_u8_smem_raw + ((1024u - (smem_u32(gemm_u8_smem_raw) & 1023u)) & 1023u));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n_tiles = (M + UBM - 1) / UBM;
  const int nu = K / 128;                           // raw stages per tile; two k-blocks each
  for (int i = threadIdx.x; i < UBN; i += UTHREADS) sm.bias[i] = bias ? __ldg(bias + i) : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < URAW; ++s) { mbar_init(&sm.raw_full[s], 1); mbar_init(&sm.raw_empty[s], 4); }
    for (int s = 0; s < UCV; ++s) { mbar_init(&sm.cv_full[s], 4); mbar_init(&sm.cv_empty[s], 1); }
    for (int s = 0; s < UWS; ++s) { mbar_init(&sm.w_full[s], 1); mbar_init(&sm.w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sm.acc_full[b], 1); mbar_init(&sm.acc_empty[b], 4); }
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
      uint32_t ur = 0, uw = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = tile * UBM;
        for (int u = 0; u < nu; ++u, ++ur) {
          const int s = ur % URAW;
          mbar_wait(&sm.raw_empty[s], ((ur / URAW) & 1u) ^ 1u);
          mbar_expect_tx(&sm.raw_full[s], (uint32_t)sizeof(sm.raw[0]));
          tma_load_2d(sm.raw[s], &tmA, &sm.raw_full[s], u * 128, m0);
          for (int half = 0; half < 2; ++half, ++uw) {
            const int sw = uw % UWS;
            mbar_wait(&sm.w_empty[sw], ((uw / UWS) & 1u) ^ 1u);
            mbar_expect_tx(&sm.w_full[sw], (uint32_t)sizeof(sm.w[0]));
            tma_load_2d(sm.w[sw], &tmW, &sm.w_full[sw], (2 * u + half) * 64, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc(FMT_F16, UBM, UBN);
    uint32_t uc = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t b = tcount & 1u;
      mbar_wait(&sm.acc_empty[b], ((tcount >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator
      tc_fence_after();
      for (int kb = 0; kb < 2 * nu; ++kb, ++uc) {
        const int sc = uc % UCV, sw = uc % UWS;                    // both rings advance once per k-block
        mbar_wait(&sm.cv_full[sc], (uc / UCV) & 1u);
        mbar_wait(&sm.w_full[sw], (uc / UWS) & 1u);
        tc_fence_after();
        const uint64_t ad = umma_desc_sw128(smem_u32(sm.cv[sc])), bd = umma_desc_sw128(smem_u32(sm.w[sw]));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_elect(tmem + b * UBN, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
        umma_commit_elect(&sm.cv_empty[sc]);
        umma_commit_elect(&sm.w_empty[sw]);
      }
      umma_commit_elect(&sm.acc_full[b]);
    }
  } else if (warp >= 8) {
    // ---- converters: thread = row of the tile
    const int r = (warp - 8) * 32 + lane, sw7 = r & 7;
    const uint32_t k1024 = 0x64006400u;                             // half2(1024, 1024)
    uint32_t ur = 0, uc = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int u = 0; u < nu; ++u, ++ur) {
        const int s = ur % URAW;
        mbar_wait(&sm.raw_full[s], (ur / URAW) & 1u);
        const uint8_t* rrow = sm.raw[s] + r * 128;
        for (int half = 0; half < 2; ++half, ++uc) {
          const int sc = uc % UCV;
          mbar_wait(&sm.cv_empty[sc], ((uc / UCV) & 1u) ^ 1u);
          unsigned char* crow = reinterpret_cast<unsigned char*>(sm.cv[sc]) + r * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 q = *reinterpret_cast<const uint4*>(rrow + (((half * 4 + c) ^ sw7) << 4));   // 16 pixels
            const uint32_t wv[4] = {q.x, q.y, q.z, q.w};
            uint32_t h[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // bytes (k0, 0x64, k1, 0x64) = half2(1024 + k0, 1024 + k1); minus 1024 is exact
              const uint32_t lo = __byte_perm(wv[e], 0x64646464u, 0x4140), hi = __byte_perm(wv[e], 0x64646464u, 0x4342);
              const __half2 a = __hsub2(*reinterpret_cast<const __half2*>(&lo), *reinterpret_cast<const __half2*>(&k1024));
              const __half2 bq = __hsub2(*reinterpret_cast<const __half2*>(&hi), *reinterpret_cast<const __half2*>(&k1024));
              h[2 * e] = *reinterpret_cast<const uint32_t*>(&a);
              h[2 * e + 1] = *reinterpret_cast<const uint32_t*>(&bq);
            }
            *reinterpret_cast<uint4*>(crow + (((2 * c) ^ sw7) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(crow + (((2 * c + 1) ^ sw7) << 4)) = make_uint4(h[4], h[5], h[6], h[7]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.cv_full[sc]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.raw_empty[s]);
      }
  } else if (warp >= 4) {
    // ---- epilogue: thread = output row
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    const float inv = __ldg(inv_scale);
    const bool vec_ok = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t b = tcount & 1u;
      mbar_wait_sleep(&sm.acc_full[b], (tcount >> 1) & 1u, 200);
      tc_fence_after();
      const int m = tile * UBM + row;
      float* crow = C + (size_t)m * ldc;
#pragma unroll 1
      for (int c = 0; c < UBN; c += 32) {
        uint32_t rr[32];
        tmem_ld32(tl + b * UBN + (uint32_t)c, rr);
        tmem_ld_wait();
        if (m < M) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = fmaf(__uint_as_float(rr[j + e]), inv, sm.bias[c + j + e]);
              if (act == ACT_RELU) x = fmaxf(x, 0.f);
              v[e] = x;
            }
            if (vec_ok) *reinterpret_cast<float4*>(crow + c + j) = make_float4(v[0], v[1], v[2], v[3]);
            else { crow[c + j] = v[0]; crow[c + j + 1] = v[1]; crow[c + j + 2] = v[2]; crow[c + j + 3] = v[3]; }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.acc_empty[b]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// |w|_max over the matrix (non-negative floats order like their bit patterns)
__global__ void absmax_kernel(const float* __restrict__ w, size_t n, uint32_t* __restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
// Wq = fp16(W s / 255) with s the power of two that puts max|Wq| in [0.5, 1); inv_scale = 1 / s
__global__ void pack_wq_kernel(const float* __restrict__ w, size_t n, const uint32_t* __restrict__ absmax_bits,
                               __half* __restrict__ out, float* __restrict__ inv_scale) {
  const float amax = __uint_as_float(*absmax_bits) / 255.f;
  int e = 0;
  if (amax > 0.f) frexpf(amax, &e);                    // amax = f 2^e, f in [0.5, 1)
  const float s = ldexpf(1.f, -e), mul = s / 255.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(w[i] * mul);
  if (blockIdx.x == 0 && threadIdx.x == 0) *inv_scale = ldexpf(1.f, e);
}

}  // namespace tc

int gemm_u8_pack_weights(const float* w, int N, int K, void* wq_f16, float* inv_scale_dev, uint32_t* scratch_dev,
                         cudaStream_t s) {
  const size_t n = (size_t)N * K;
  AVHOT_CUDA(cudaMemsetAsync(scratch_dev, 0, sizeof(uint32_t), s));
  tc::absmax_kernel<<<num_sms() * 4, 256, 0, s>>>(w, n, scratch_dev);
  AVHOT_LAUNCH_CHECK();
  tc::pack_wq_kernel<<<num_sms() * 4, 256, 0, s>>>(w, n, scratch_dev, reinterpret_cast<__half*>(wq_f16), inv_scale_dev);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

bool gemm_u8_tc_supported(const uint8_t* a, int lda, int N, int K) {
  return N == tc::UBN && K > 0 && (K % 128) == 0 && (lda % 16) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0;
}

int gemm_u8_tc(const uint8_t* a, int lda, const void* wq_f16, const float* bias, const float* inv_scale_dev, float* c,
               int ldc, int M, int N, int K, int act, cudaStream_t s, const char* prof_name) {
  using namespace tc;
  AVHOT_REQUIRE(a && wq_f16 && inv_scale_dev && c && M >= 0, AVHOT_ERR_INVALID, "gemm_u8_tc: bad arguments");
  AVHOT_REQUIRE(gemm_u8_tc_supported(a, lda, N, K), AVHOT_ERR_UNSUPPORTED, "gemm_u8_tc: N must be 256, K a multiple of 128");
  if (M == 0) return AVHOT_OK;
  CUtensorMap tmA, tmW;
  int rc = make_tensor_map_2d(&tmA, a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, (uint64_t)M, (uint64_t)K, (uint64_t)lda, UBM, 128);
  if (rc != AVHOT_OK) return rc;
  rc = make_tensor_map_2d(&tmW, wq_f16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)N, (uint64_t)K, (uint64_t)K, UBN, 64);
  if (rc != AVHOT_OK) return rc;
  const int smem = (int)sizeof(GemmU8Smem) + 1024;
  AVHOT_CUDA(cudaFuncSetAttribute(gemm_u8_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_tiles = ceil_div(M, UBM);
  ProfScope prof(prof_name, s);
  gemm_u8_f16_kernel<<<min(n_tiles, num_sms()), UTHREADS, smem, s>>>(tmA, tmW, bias, inv_scale_dev, c, ldc, M, K, act);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

}  // namespace avhot
