// FrameEncoder's Linear + ReLU on fp32 frames - and, with U8 = true, on uint8 pixels - (AVHOT_PREC_TC path of a4,
// src/encoders.py:225-229) as an fp16 tensor-core GEMM with the conversion inside the kernel, on CTA pairs:
//     C[M, 256] = relu( (fp16(A[M, K]) x Wq[256, K]^T) * c / s + bias ),   Wq = fp16(W s / 255) (gemm_u8.cu), c = 255
// Why not kind::tf32 on the fp32 operands (gemm_tc.cu): that kernel streams the fp32 WEIGHTS next to the frames - per
// 64-column k-block 64 KB of weights from L2 for 32 KB of frames from HBM, 45 B/clk per SM, the ingest limit of an SM -
// and so sits at 0.63-0.67 of the HBM roofline although DRAM traffic is the algorithmic minimum.  Here the weights are
// fp16 (32 KB per k-block) and a CTA pair (cta_group::2, M = 256) shares them: each CTA loads 128 of the 256 weight
// rows, 16 KB per k-block against its 32 KB of frames.  fp16 keeps 11 significant bits of a frame value (TF32: 10) and
// runs at twice the TF32 rate.  Frame values must be fp16-representable in magnitude (|x| < 65504; pixels / 255 are).
// uint8 pixels (c = 1; exact in fp16, PRMT + HSUB2 as in gemm_u8.cu, whose single-CTA kernel was bound by the same
// weight stream: 32 KB of weights per 8 KB of pixels) take the same pair structure: a raw stage is then two k-blocks.
//
// Persistent pair (one CTA per SM, grid-stride over 256-row tiles), 12 warps per CTA:
//   warp 0      TMA producer: fp32 boxes (128 rows x 32 floats, 128-byte swizzle), two per k-block, 3-stage ring; this CTA's
//               half of the fp16 weight k-block (128 rows x 64 k) into a 4-stage ring, completing the LEADER's barrier;
//   warps 8-11  converters, thread = row: 16 x 128-bit shared loads, cvt.rn.f16x2.f32, 8 x 128-bit stores of the A operand's
//               K-major swizzled k-block (3-stage ring); arrive on the leader's cv_full;
//   warp 1      leader CTA: one elected lane issues tcgen05.mma.cta_group::2 kind::f16 (M 256, N 256, K 16) x 4 per k-block
//               into one of two 256-column TMEM accumulators; commits (multicast) free both CTAs' stages;
//   warps 4-7   epilogue of the previous tile while the next one accumulates (thread = row): scale, bias, ReLU, stores.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace avhot {
namespace tc {

constexpr int VBM = 128, VBN = 256, VRAW = 3, VCV = 3, VWS = 4, VTHREADS = 384;

struct GemmCvtSmem {
  float raw[VRAW][2][VBM * 32];     // fp32: a k-block as two TMA boxes (columns 0-31, 32-63); u8: raw[s][0] = one box of
                                    // 128 rows x 128 pixels = two k-blocks
  __half cv[VCV][VBM * 64];         // converted A k-block (this CTA's 128 rows)
  __half w[VWS][(VBN / 2) * 64];    // this CTA's 128 weight rows of a k-block
  float bias[VBN];
  uint64_t raw_full[VRAW], raw_empty[VRAW], cv_full[VCV], cv_empty[VCV], w_full[VWS], w_empty[VWS], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(GemmCvtSmem) + 1024 <= 227 * 1024, "one CTA per SM");

template <bool U8>
__global__ void __launch_bounds__(VTHREADS, 1)
gemm_cvt_f16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                         const float* __restrict__ bias, const float* __restrict__ inv_scale, float scale_mul,
                         float* __restrict__ C, int ldc, int M, int K, int act) {
  extern __shared__ unsigned char gemm_cvt_smem_raw[];
  GemmCvtSmem& sm = *reinterpret_cast<GemmCvtSmem*>(gemm_cvt_smem_raw + ((1024u - (smem_u32(gemm_cvt_smem_raw) & 1023u)) & 1023u));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int n_ptiles = (M + 2 * VBM - 1) / (2 * VBM);        // 256-row tiles, one per pair and iteration
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int nkb = K / 64;
  for (int i = threadIdx.x; i < VBN; i += VTHREADS) sm.bias[i] = bias ? __ldg(bias + i) : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < VRAW; ++s) { mbar_init(&sm.raw_full[s], 1); mbar_init(&sm.raw_empty[s], 4); }
    for (int s = 0; s < VCV; ++s) { mbar_init(&sm.cv_full[s], 4 * 2); mbar_init(&sm.cv_empty[s], 1); }   // both CTAs' converters
    for (int s = 0; s < VWS; ++s) { mbar_init(&sm.w_full[s], 1); mbar_init(&sm.w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sm.acc_full[b], 1); mbar_init(&sm.acc_empty[b], 4 * 2); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_pair(&sm.tmem_base, 512); tmem_relinquish_pair(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t wfull0 = cluster_map_shared(&sm.w_full[0], 0);
      uint32_t ur = 0, uw = 0;
      for (int pt = pair; pt < n_ptiles; pt += n_pairs) {
        const int m0 = pt * 2 * VBM + (int)crank * VBM;
        for (int kb = 0; kb < nkb; ++kb, ++uw) {
          if (!U8 || (kb & 1) == 0) {
            const int s = ur % VRAW;
            mbar_wait(&sm.raw_empty[s], ((ur / VRAW) & 1u) ^ 1u);
            if (U8) {
              mbar_expect_tx(&sm.raw_full[s], (uint32_t)sizeof(sm.raw[0][0]));
              tma_load_2d(sm.raw[s][0], &tmA, &sm.raw_full[s], kb * 64, m0);        // 128 pixels = k-blocks kb, kb + 1
            } else {
              mbar_expect_tx(&sm.raw_full[s], (uint32_t)sizeof(sm.raw[0]));
              tma_load_2d(sm.raw[s][0], &tmA, &sm.raw_full[s], kb * 64, m0);
              tma_load_2d(sm.raw[s][1], &tmA, &sm.raw_full[s], kb * 64 + 32, m0);
            }
            ++ur;
          }
          const int sw = uw % VWS;
          mbar_wait(&sm.w_empty[sw], ((uw / VWS) & 1u) ^ 1u);
          if (leader) mbar_expect_tx(&sm.w_full[sw], 2u * (uint32_t)sizeof(sm.w[0]));
          tma_load_2d_pair(sm.w[sw], &tmW, wfull0 + (uint32_t)sw * 8u, kb * 64, (int)crank * (VBN / 2));
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = umma_idesc(FMT_F16, 2 * VBM, VBN);
      uint32_t uc = 0, tcount = 0;
      for (int pt = pair; pt < n_ptiles; pt += n_pairs, ++tcount) {
        const uint32_t b = tcount & 1u;
        mbar_wait_cluster(&sm.acc_empty[b], ((tcount >> 1) & 1u) ^ 1u);      // both CTAs' epilogues drained this accumulator
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++uc) {
          const int sc = uc % VCV, sw = uc % VWS;
          mbar_wait_cluster(&sm.cv_full[sc], (uc / VCV) & 1u);
          mbar_wait(&sm.w_full[sw], (uc / VWS) & 1u);
          tc_fence_after();
          const uint64_t ad = umma_desc_sw128(smem_u32(sm.cv[sc])), bd = umma_desc_sw128(smem_u32(sm.w[sw]));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_pair_elect(tmem + b * VBN, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair_mc_elect(&sm.cv_empty[sc], 3);
          umma_commit_pair_mc_elect(&sm.w_empty[sw], 3);
        }
        umma_commit_pair_mc_elect(&sm.acc_full[b], 3);
      }
    }
  } else if (warp >= 8) {
    // ---- converters: thread = row of this CTA's tile
    const int r = (warp - 8) * 32 + lane, sw7 = r & 7;
    const uint32_t cv_full0 = cluster_map_shared(&sm.cv_full[0], 0);
    uint32_t ur = 0, uc = 0;
    const uint32_t k1024 = 0x64006400u;               // half2(1024, 1024)
    for (int pt = pair; pt < n_ptiles; pt += n_pairs)
      for (int kb = 0; kb < nkb; ++kb, ++uc) {
        const int s = ur % VRAW, sc = uc % VCV;
        if (!U8 || (kb & 1) == 0) mbar_wait(&sm.raw_full[s], (ur / VRAW) & 1u);
        mbar_wait(&sm.cv_empty[sc], ((uc / VCV) & 1u) ^ 1u);
        unsigned char* crow = reinterpret_cast<unsigned char*>(sm.cv[sc]) + r * 128;
        if (U8) {
          const unsigned char* rrow = reinterpret_cast<const unsigned char*>(sm.raw[s][0]) + r * 128;
          const int half = kb & 1;
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
        } else {
#pragma unroll
          for (int bx = 0; bx < 2; ++bx) {
            const unsigned char* rrow = reinterpret_cast<const unsigned char*>(sm.raw[s][bx]) + r * 128;
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
              const float4 f0 = *reinterpret_cast<const float4*>(rrow + ((c ^ sw7) << 4));
              const float4 f1 = *reinterpret_cast<const float4*>(rrow + (((c + 1) ^ sw7) << 4));
              const __half2 h0 = __floats2half2_rn(f0.x, f0.y), h1 = __floats2half2_rn(f0.z, f0.w);
              const __half2 h2 = __floats2half2_rn(f1.x, f1.y), h3 = __floats2half2_rn(f1.z, f1.w);
              // floats bx*32 + 4c .. + 7 of the k-block = its 16-byte fp16 chunk bx*4 + c/2
              *reinterpret_cast<uint4*>(crow + (((bx * 4 + (c >> 1)) ^ sw7) << 4)) =
                  make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                             *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        const bool raw_done = !U8 || (kb & 1) == 1 || kb == nkb - 1;
        if (lane == 0) {
          mbar_arrive_cluster(cv_full0 + (uint32_t)sc * 8u);
          if (raw_done) mbar_arrive(&sm.raw_empty[s]);
        }
        if (raw_done) ++ur;
      }
  } else if (warp >= 4) {
    // ---- epilogue: thread = output row of this CTA's half of the tile
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    const float inv = __ldg(inv_scale) * scale_mul;
    const bool vec_ok = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
    const uint32_t acc_empty0 = cluster_map_shared(&sm.acc_empty[0], 0);
    uint32_t tcount = 0;
    for (int pt = pair; pt < n_ptiles; pt += n_pairs, ++tcount) {
      const uint32_t b = tcount & 1u;
      mbar_wait_sleep(&sm.acc_full[b], (tcount >> 1) & 1u, 200);
      tc_fence_after();
      const int m = pt * 2 * VBM + (int)crank * VBM + row;
      float* crow = C + (size_t)m * ldc;
#pragma unroll 1
      for (int c = 0; c < VBN; c += 32) {
        uint32_t rr[32];
        tmem_ld32(tl + b * VBN + (uint32_t)c, rr);
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
      if (lane == 0) mbar_arrive_cluster(acc_empty0 + b * 8u);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem, 512);
}

}  // namespace tc

bool gemm_u8_pair_tc_supported(const uint8_t* a, int lda, int N, int K) {
  static const bool off = [] { const char* e = getenv("AVHOT_FRAME_U8_PAIR"); return e && atoi(e) == 0; }();
  return !off && N == tc::VBN && K > 0 && (K % 128) == 0 && (lda % 16) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0;
}

bool gemm_f32cvt_tc_supported(const float* a, int lda, int N, int K) {
  static const bool off = [] { const char* e = getenv("AVHOT_FRAME_CVT"); return e && atoi(e) == 0; }();
  return !off && N == tc::VBN && K > 0 && (K % 64) == 0 && (lda % 4) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0;
}

// wq_f16 / inv_scale_dev: gemm_u8_pack_weights (Wq = fp16(W s / 255)); scale_mul = 255 undoes the pixel scaling for fp32
// frames; a_u8 != NULL: uint8 pixels (scale_mul = 1)
static int gemm_cvt_launch(const float* a_f32, const uint8_t* a_u8, int lda, const void* wq_f16, const float* bias,
                           const float* inv_scale_dev, float scale_mul, float* c, int ldc, int M, int N, int K, int act,
                           cudaStream_t s, const char* prof_name) {
  using namespace tc;
  if (M == 0) return AVHOT_OK;
  CUtensorMap tmA, tmW;
  int rc = a_u8 ? make_tensor_map_2d(&tmA, a_u8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, (uint64_t)M, (uint64_t)K, (uint64_t)lda, VBM, 128)
                : make_tensor_map_2d(&tmA, a_f32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)M, (uint64_t)K, (uint64_t)lda, VBM, 32);
  if (rc != AVHOT_OK) return rc;
  rc = make_tensor_map_2d(&tmW, wq_f16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)N, (uint64_t)K, (uint64_t)K, VBN / 2, 64);
  if (rc != AVHOT_OK) return rc;
  const int smem = (int)sizeof(GemmCvtSmem) + 1024;
  auto kern = a_u8 ? gemm_cvt_f16_pair_kernel<true> : gemm_cvt_f16_pair_kernel<false>;
  AVHOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_ptiles = ceil_div(M, 2 * VBM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * min(n_ptiles, num_sms() / 2)));
  cfg.blockDim = dim3(VTHREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof(prof_name, s);
  AVHOT_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmW, bias, inv_scale_dev, scale_mul, c, ldc, M, K, act));
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

int gemm_f32cvt_tc(const float* a, int lda, const void* wq_f16, const float* bias, const float* inv_scale_dev, float scale_mul,
                   float* c, int ldc, int M, int N, int K, int act, cudaStream_t s, const char* prof_name) {
  AVHOT_REQUIRE(a && wq_f16 && inv_scale_dev && c && M >= 0, AVHOT_ERR_INVALID, "gemm_f32cvt_tc: bad arguments");
  AVHOT_REQUIRE(gemm_f32cvt_tc_supported(a, lda, N, K), AVHOT_ERR_UNSUPPORTED, "gemm_f32cvt_tc: N must be 256, K a multiple of 64");
  return gemm_cvt_launch(a, nullptr, lda, wq_f16, bias, inv_scale_dev, scale_mul, c, ldc, M, N, K, act, s, prof_name);
}

int gemm_u8_pair_tc(const uint8_t* a, int lda, const void* wq_f16, const float* bias, const float* inv_scale_dev, float* c,
                    int ldc, int M, int N, int K, int act, cudaStream_t s, const char* prof_name) {
  AVHOT_REQUIRE(a && wq_f16 && inv_scale_dev && c && M >= 0, AVHOT_ERR_INVALID, "gemm_u8_pair_tc: bad arguments");
  AVHOT_REQUIRE(gemm_u8_pair_tc_supported(a, lda, N, K), AVHOT_ERR_UNSUPPORTED, "gemm_u8_pair_tc: N must be 256, K a multiple of 128");
  return gemm_cvt_launch(nullptr, a, lda, wq_f16, bias, inv_scale_dev, 1.f, c, ldc, M, N, K, act, s, prof_name);
}

}  // namespace avhot
