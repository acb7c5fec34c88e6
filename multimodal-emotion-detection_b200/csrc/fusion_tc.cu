// HybridFusion per-pass work on the tcgen05 tensor cores (AVHOT_PREC_TC path of a5 + a6, src/fusion.py:455-530,
// src/attention.py:80-141), for rows (clip, pass) whose audio token is pass-invariant (fusion_clip_kernel has
// already produced Z_a, Q_0, K_0a, V_0a, K_1a, V_1a per clip).
//
// One CTA = 128 rows.  Thread r of the four epilogue warps owns row r = TMEM lane r, so every non-GEMM step of the
// module - both LayerNorms, the 4-head / 2-key softmax, masks, Philox dropout, residuals, the gate's ReLU + dropout
// + 128-wide dot, the 8-class classifier - is thread-local arithmetic over the 256 accumulator columns of its
// lane: no shuffles, no shared-memory reductions.  The eight GEMM groups of a tile
//   G1 proj_v              x = z_v        -> Z_v raw            G5 Q_1 | K_1v          x^ -> D_a | D_b
//   G2 K_0v | V_0v         x^             -> D_a | D_b          G6 V_1v                x^ -> D_a
//   G3 out_proj_0          ctx_0          -> D_a                G7 out_proj_1          ctx_1 -> D_b
//   G4 gate(H_0)           H_0            -> D_a[0:128]         G8 gate(H_1)           H_1 -> D_a[0:128]
// run as tcgen05.mma kind::f16 (M 128, N 128, K 16) with fp16 A operands that the epilogue writes straight into
// the 128-byte-swizzled K-major layout (two 64 KB buffers X and C), fp16 weights streamed by TMA through a
// 4-stage ring, and two 256-column fp32 accumulators in TMEM.  The classifier is linear, so
// logits = w_0 cls(H_0) + w_1 cls(H_1) + b and H_m never has to outlive its own epilogue.
//   warp 0  TMA producer      warp 1  TMEM alloc + MMA issue      warps 2-5 (6-9)  epilogue (thread = row)
// NH = 2 (default): eight epilogue warps.  Warps w and w + 4 own the same 32 rows (the same TMEM lane quarter) and split
// the accumulator COLUMNS in halves - heads {0, 1} / {2, 3}, 128 of the 256 LayerNorm columns, 64 of the 128 gate
// units - exchanging only the LayerNorm partial sums (through a 4 KB shared-memory mailbox and a 64-thread named
// barrier per warp pair) and, once per tile, the gate / classifier partial sums.  With one epilogue warp per scheduler
// (NH = 1) the kernel issued 0.24 instructions per cycle per warp: every dependent instruction exposed its full
// latency; two warps per scheduler nearly double the epilogue rate.
#include <stdio.h>
#include <stdlib.h>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tc.cuh"
#include "fusion_tc.cuh"

namespace avhot {
namespace tc {

constexpr int UB = 128, UD = 256, UST = 4;
// rows of the packed fp16 weight tensor (K = 256 columns each)
constexpr int ROW_PROJ = 0, ROW_K0 = 256, ROW_V0 = 512, ROW_Q1 = 768, ROW_K1 = 1024, ROW_V1 = 1280, ROW_O0 = 1536,
              ROW_O1 = 1792, ROW_G0 = 2048, ROWS_ALL = 2176;

struct FusTcSmem {
  __half X[4][UB * 64];      // A operand: z_v, then x^_v
  __half C[4][UB * 64];      // A operand: ctx_0 -> H_0 -> ctx_1 -> H_1
  __half wB[UST][128 * 64];  // weight ring, 128 output rows x 64 k per stage
  float vec[FUS_VEC_FLOATS];
  float2 xch[2][4][UB];      // LayerNorm partial sums: [parity of the exchange][column part][row]
  uint64_t full[UST], empty[UST], m2e, e2m;
  uint32_t tmem_base;
};

struct SubGemm { int row0, n, from_c, dcol; };
struct Group { int n_sub; SubGemm sub[2]; };
__device__ __constant__ Group kGroups[8] = {
    {1, {{ROW_PROJ, 256, 0, 0}, {0, 0, 0, 0}}},
    {2, {{ROW_K0, 256, 0, 0}, {ROW_V0, 256, 0, 256}}},
    {1, {{ROW_O0, 256, 1, 0}, {0, 0, 0, 0}}},
    {1, {{ROW_G0, 128, 1, 0}, {0, 0, 0, 0}}},
    {2, {{ROW_Q1, 256, 0, 0}, {ROW_K1, 256, 0, 256}}},
    {1, {{ROW_V1, 256, 0, 0}, {0, 0, 0, 0}}},
    {1, {{ROW_O1, 256, 1, 256}, {0, 0, 0, 0}}},
    {1, {{ROW_G0, 128, 1, 0}, {0, 0, 0, 0}}},
};

// 8 consecutive fp16 of row r starting at column c (c % 8 == 0) in a [4][128 x 64] swizzled K-major operand
__device__ __forceinline__ uint4* a_chunk(__half* buf, int r, int c) {
  unsigned char* p = reinterpret_cast<unsigned char*>(buf) + (size_t)(c >> 6) * (UB * 128) + (size_t)r * 128 +
                     ((((c & 63) >> 3) ^ (r & 7)) << 4);
  return reinterpret_cast<uint4*>(p);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint32_t pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 h2 = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
  }
  return make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

template <int NH>
__global__ void __launch_bounds__(64 + 128 * NH, 1)
fusion_tc_kernel(const __grid_constant__ CUtensorMap tmW, const FusTcParams P) {
  constexpr int UTHREADS = 64 + 128 * NH;
  constexpr int HC = UD / NH;                           // accumulator columns per epilogue thread
  extern __shared__ unsigned char fus_tc_smem_raw[];
  FusTcSmem& sm = *reinterpret_cast<FusTcSmem*>(fus_tc_smem_raw + ((1024u - (smem_u32(fus_tc_smem_raw) & 1023u)) & 1023u));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const long long n_rows = (long long)P.B * P.T;
  const int n_tiles = (int)((n_rows + UB - 1) / UB);
  const int kv = P.dim_v / 64;                          // k-blocks of the video embedding

  for (int i = threadIdx.x; i < FUS_VEC_FLOATS; i += UTHREADS) sm.vec[i] = P.vec[i];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < UST; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.m2e, 1);
    mbar_init(&sm.e2m, 4 * NH);                         // one arrival per epilogue warp
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
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int g = 0; g < 8; ++g)
          for (int sb = 0; sb < kGroups[g].n_sub; ++sb) {
            const SubGemm sg = kGroups[g].sub[sb];
            const int nkb = g == 0 ? kv : 4;
            for (int nh = 0; nh < sg.n / 128; ++nh)
              for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % UST;
                mbar_wait(&sm.empty[s], ((it / UST) & 1) ^ 1);
                mbar_expect_tx(&sm.full[s], (uint32_t)sizeof(sm.wB[0]));
                tma_load_2d(sm.wB[s], &tmW, &sm.full[s], kb * 64, sg.row0 + nh * 128);
              }
          }
    }
  } else if (warp == 1) {
    {   // the whole warp, converged: one elected lane issues (tc.cuh, umma_*_elect)
      constexpr uint32_t idesc = umma_idesc(FMT_F16, UB, 128);
      uint32_t it = 0, op = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int g = 0; g < 8; ++g, ++op) {
          mbar_wait(&sm.e2m, op & 1);                   // the epilogue has produced this group's A / freed its D
          tc_fence_after();
          for (int sb = 0; sb < kGroups[g].n_sub; ++sb) {
            const SubGemm sg = kGroups[g].sub[sb];
            const int nkb = g == 0 ? kv : 4;
            __half* abuf = sg.from_c ? &sm.C[0][0] : &sm.X[0][0];
            for (int nh = 0; nh < sg.n / 128; ++nh)
              for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % UST;
                mbar_wait(&sm.full[s], (it / UST) & 1);
                tc_fence_after();
                const uint64_t adesc = umma_desc_sw128(smem_u32(abuf + (size_t)kb * (UB * 64)));
                const uint64_t bdesc = umma_desc_sw128(smem_u32(sm.wB[s]));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_elect(tmem + (uint32_t)(sg.dcol + nh * 128), adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                umma_commit_elect(&sm.empty[s]);
              }
          }
          umma_commit_elect(&sm.m2e);
        }
    }
  } else {
    const int quarter = warp & 3;
    const int half = NH == 1 ? 0 : (warp - 2) >> 2;     // column half of this warp (warps w and w + 4 share a quarter)
    const int cb = half * HC;                           // first accumulator column of this thread
    const int r = quarter * 32 + lane;
    uint32_t xn_ = 0;                                   // exchanges done so far (mailbox parity)
    // total of a per-row partial (sum, sum of squares) over the column halves
    auto row_total = [&](float& a, float& b) {
      if (NH == 1) return;
      sm.xch[xn_ & 1][half][r] = make_float2(a, b);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(32 * NH) : "memory");
      // the same order of addition in every part, so all see bit-identical statistics
      float2 t = sm.xch[xn_ & 1][0][r];
#pragma unroll
      for (int h2 = 1; h2 < NH; ++h2) { const float2 o = sm.xch[xn_ & 1][h2][r]; t.x += o.x; t.y += o.y; }
      a = t.x; b = t.y;
      ++xn_;
    };
    const uint32_t tl = tmem + ((uint32_t)(quarter * 32) << 16);
    const float* V = sm.vec;
    uint32_t op = 0;
    auto to_mma = [&]() {           // A operand written / accumulators drained: let the MMA warp start the next group
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.e2m);
    };
    auto from_mma = [&]() {
      mbar_wait(&sm.m2e, op & 1);
      ++op;
      tc_fence_after();
    };
    // debug trace: the first epilogue thread of CTA 0 stamps clock64() at every phase boundary of its third tile
    const bool probe = P.trace && blockIdx.x == 0 && warp == 2 && lane == 0;
    int tile_no = 0, ev = 0;
    auto TR = [&]() { if (probe && tile_no == 2 && ev < 64) P.trace[ev++] = clock64(); };
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_no) {
      long long row = (long long)tile * UB + r;
      TR();                                                               // 0: tile start
      const bool live = row < n_rows;
      if (!live) row = n_rows - 1;
      const int b = (int)((uint32_t)row / (uint32_t)P.T), t = (int)((uint32_t)row - (uint32_t)b * (uint32_t)P.T);
      const uint32_t gclip = ds_clip(P.ds, b), pass = P.ds.t0 + (uint32_t)t;
      const float* cbuf = P.clipbuf + (size_t)b * (6 * UD);
      const bool inv0 = P.mask ? (__ldg(P.mask + (size_t)b * 2) <= 0.f) : false;
      const bool inv1 = P.mask ? (__ldg(P.mask + (size_t)b * 2 + 1) <= 0.f) : false;

      // ---- E0: the tile's z_v rows -> fp16 A operand ----------------------------------------------------------------
      // A cooperative, coalesced copy: the 4 NH epilogue warps take the tile's rows round robin, lane l converts floats
      // [8 l, 8 l + 8) of a row (a warp reads a whole 1 KB row with two instructions) and writes the 16-byte chunk where
      // the UMMA layout wants it.  (Thread = row loads - 32 different lines per instruction - cost 11 k of a tile's 124 k
      // cycles in the r4l clock trace.)  Only the MMA reads these rows, and it waits for every warp's to_mma().
      {
        const int ew = warp - 2;
        if (8 * lane < P.dim_v) {
#pragma unroll 4
          for (int rr = ew; rr < UB; rr += 4 * NH) {
            // 32-bit row arithmetic (fusion_tc_forward checks B * T < 2^31): a 64-bit division per row cost more than
            // the loads themselves
            uint32_t grow = (uint32_t)tile * UB + (uint32_t)rr;
            if (grow >= (uint32_t)n_rows) grow = (uint32_t)n_rows - 1u;
            const uint32_t gb = grow / (uint32_t)P.T, gt = grow - gb * (uint32_t)P.T;
            const float* zr = P.z_v + ((size_t)gb * P.rpc_v + (P.rpc_v == 1 ? 0u : gt)) * P.dim_v + 8 * lane;
            // the same row of this CTA's next tile -> L2 (rows of consecutive (clip, pass) are consecutive in memory);
            // one request per 128-byte line
            if (P.rpc_v != 1 && (lane & 3) == 0 && (long long)grow + (long long)gridDim.x * UB < n_rows)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(zr + (size_t)gridDim.x * UB * P.dim_v));
            const float4 u0 = __ldg(reinterpret_cast<const float4*>(zr));
            const float4 u1 = __ldg(reinterpret_cast<const float4*>(zr + 4));
            const float v[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
            *a_chunk(&sm.X[0][0], rr, 8 * lane) = pack8(v);
          }
        }
      }
      TR();                                                               // 1: E0 done
      to_mma();                                                           // -> G1

      // ---- E1: Z_v = pre_ln(proj), x^_v = (Z_v - mean) / std -> X ----------------------------------------------------
      from_mma();
      TR();                                                               // 2: G1 arrived
      float zmean, zstd;
      {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = cb; c < cb + HC; c += 32) {
          uint32_t a[32];
          tmem_ld32(tl + (uint32_t)c, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { const float x = __uint_as_float(a[j]) + V[VEC_PROJ_B + c + j]; s1 += x; s2 = fmaf(x, x, s2); }
        }
        row_total(s1, s2);
        const float m1 = s1 * (1.f / UD);
        const float rstd1 = 1.f / sqrtf(fmaxf(s2 * (1.f / UD) - m1 * m1, 0.f) + 1e-5f);
        s1 = 0.f; s2 = 0.f;
#pragma unroll 1
        for (int c = cb; c < cb + HC; c += 32) {
          uint32_t a[32];
          tmem_ld32(tl + (uint32_t)c, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(a[j]) + V[VEC_PROJ_B + c + j];
            const float z = (x - m1) * rstd1 * V[VEC_PRE_W + c + j] + V[VEC_PRE_B + c + j];
            s1 += z; s2 = fmaf(z, z, s2);
          }
        }
        row_total(s1, s2);
        zmean = s1 * (1.f / UD);
        zstd = sqrtf(fmaxf(s2 * (1.f / UD) - zmean * zmean, 0.f) + 1e-5f);
        const float rstd2 = 1.f / zstd;
#pragma unroll 1
        for (int c = cb; c < cb + HC; c += 32) {
          uint32_t a[32];
          tmem_ld32(tl + (uint32_t)c, a);
          tmem_ld_wait();
          float xn[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(a[j]) + V[VEC_PROJ_B + c + j];
            const float z = (x - m1) * rstd1 * V[VEC_PRE_W + c + j] + V[VEC_PRE_B + c + j];
            xn[j] = (z - zmean) * rstd2;
          }
#pragma unroll
          for (int j = 0; j < 32; j += 8) *a_chunk(&sm.X[0][0], r, c + j) = pack8(&xn[j]);
        }
      }
      TR();                                                               // 3: E1 done
      to_mma();                                                           // -> G2 (K_0v | V_0v)

      float cls0[8], cls1[8], score0 = 0.f, score1 = 0.f;
      // ---- the two attention layers --------------------------------------------------------------------------------
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        constexpr int HPT = 4 / NH;                                       // heads per thread: heads hb .. hb + HPT - 1
        const int hb = half * HPT;
        float aa[HPT], av[HPT];                                           // attention weights per head: audio / video key
        if (layer == 0) {
          from_mma();                                                     // D_a = K_0v, D_b = V_0v
          TR();                                                           // 4: G2 arrived
          // scores: query = audio token (per clip), keys = audio token (per clip) and video token (this row)
#pragma unroll 1
          for (int hh = 0; hh < HPT; ++hh) {
            float sa = 0.f, sv = 0.f;
#pragma unroll 1
            for (int c = (hb + hh) * 64; c < (hb + hh) * 64 + 64; c += 32) {
              uint32_t a[32];
              tmem_ld32(tl + (uint32_t)c, a);
              tmem_ld_wait();
              // (the per-clip vectors are 16-byte aligned: 128-bit loads - as scalars they were 195 LDG sites, most of the
              // kernel's long-scoreboard stalls with two warps per scheduler)
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 q4 = __ldg(reinterpret_cast<const float4*>(cbuf + 1 * UD + c + j));
                const float4 k4 = __ldg(reinterpret_cast<const float4*>(cbuf + 2 * UD + c + j));
                const float q[4] = {q4.x, q4.y, q4.z, q4.w}, kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  sa = fmaf(q[e], kk[e], sa);
                  sv = fmaf(q[e], __uint_as_float(a[j + e]) + V[VEC_BK0 + c + j + e], sv);
                }
              }
            }
            aa[hh] = sa; av[hh] = sv;
          }
        } else {
          from_mma();                                                     // D_a = Q_1, D_b = K_1v
          TR();                                                           // 12: G5 arrived
#pragma unroll 1
          for (int hh = 0; hh < HPT; ++hh) {
            float sa = 0.f, sv = 0.f;
#pragma unroll 1
            for (int c = (hb + hh) * 64; c < (hb + hh) * 64 + 64; c += 32) {
              uint32_t a[32], k[32];
              tmem_ld32(tl + (uint32_t)c, a);
              tmem_ld32(tl + 256u + (uint32_t)c, k);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 k4 = __ldg(reinterpret_cast<const float4*>(cbuf + 4 * UD + c + j));
                const float ka[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float q = __uint_as_float(a[j + e]) + V[VEC_BQ1 + c + j + e];
                  sa = fmaf(q, ka[e], sa);
                  sv = fmaf(q, __uint_as_float(k[j + e]) + V[VEC_BK1 + c + j + e], sv);
                }
              }
            }
            aa[hh] = sa; av[hh] = sv;
          }
        }
        // softmax over the two keys, masks, dropout (src/attention.py:112-130)
        {
          // element (head, key) = 2 head + key of the (4, 1, 2) mask: block 0 = heads 0, 1; block 1 = heads 2, 3
          const Philox4 w0 = philox4x32_10_rk((uint32_t)(hb >> 1), gclip, pass, AVHOT_SITE_ATTN + layer, P.ds);
          Philox4 w1 = w0;
          if (NH == 1) w1 = philox4x32_10_rk(1u, gclip, pass, AVHOT_SITE_ATTN + layer, P.ds);
#pragma unroll
          for (int hh = 0; hh < HPT; ++hh) {
            float s0 = aa[hh] * 0.125f, s1 = av[hh] * 0.125f;
            if (inv0) s0 = -1e4f;
            if (inv1) s1 = -1e4f;
            const float mx = fmaxf(s0, s1);
            const float e0 = __expf(s0 - mx), e1 = __expf(s1 - mx);
            const float den = e0 + e1;
            float a0 = e0 / den, a1 = e1 / den;
            if (inv0 && inv1) { a0 = 0.f; a1 = 0.f; }
            if (P.ds.thr != 0u) {
              const Philox4& w = hh < 2 ? w0 : w1;
              a0 = philox_word(w, (uint32_t)(((hb + hh) & 1) * 2)) >= P.ds.thr ? a0 * P.ds.scale : 0.f;
              a1 = philox_word(w, (uint32_t)(((hb + hh) & 1) * 2 + 1)) >= P.ds.thr ? a1 * P.ds.scale : 0.f;
            }
            aa[hh] = a0; av[hh] = a1;
          }
        }
        TR();                                                             // 5 / 13: scores + softmax done
        if (layer == 1) {
          to_mma();                                                       // D_a, D_b drained -> G6 (V_1v -> D_a)
          from_mma();
          TR();                                                           // 14: G6 arrived
        }
        // context = a_a V_a (per clip) + a_v V_v (this row) -> C     (layer 0: V_0v is in D_b; layer 1: V_1v in D_a)
        {
          const uint32_t vcol = layer == 0 ? 256u : 0u;
          const int vb = layer == 0 ? VEC_BV0 : VEC_BV1, vslot = layer == 0 ? 3 : 5;
#pragma unroll 1
          for (int c = cb; c < cb + HC; c += 32) {
            uint32_t a[32];
            tmem_ld32(tl + vcol + (uint32_t)c, a);
            tmem_ld_wait();
            const int hh = (c >> 6) - hb;
            float ctx[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 v4 = __ldg(reinterpret_cast<const float4*>(cbuf + vslot * UD + c + j));
              const float va[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                ctx[j + e] = aa[hh] * va[e] + av[hh] * (__uint_as_float(a[j + e]) + V[vb + c + j + e]);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 8) *a_chunk(&sm.C[0][0], r, c + j) = pack8(&ctx[j]);
          }
        }
        TR();                                                             // 6 / 15: context written
        to_mma();                                                         // -> G3 / G7 (out_proj)

        // out_proj + residual -> post_ln -> H: classifier partial sums, fp16 H -> C for the gate GEMM
        from_mma();
        TR();                                                             // 7 / 16: G3 / G7 arrived
        {
          const uint32_t ocol = layer == 0 ? 0u : 256u;
          const int ob = layer == 0 ? VEC_BO0 : VEC_BO1;
          float s1 = 0.f, s2 = 0.f;
          auto residual = [&](int c, float (&res)[32]) {
            if (layer == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {                                             // Z_a
                const float4 z4 = __ldg(reinterpret_cast<const float4*>(cbuf + c + j));
                res[j] = z4.x; res[j + 1] = z4.y; res[j + 2] = z4.z; res[j + 3] = z4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {                                             // Z_v = x^ * std + mean
                const uint4 u = *a_chunk(&sm.X[0][0], r, c + j);
                const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 f = __half22float2(h2[q]);
                  res[j + 2 * q] = fmaf(f.x, zstd, zmean);
                  res[j + 2 * q + 1] = fmaf(f.y, zstd, zmean);
                }
              }
            }
          };
#pragma unroll 1
          for (int c = cb; c < cb + HC; c += 32) {
            uint32_t a[32];
            float res[32];
            tmem_ld32(tl + ocol + (uint32_t)c, a);
            residual(c, res);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float x = __uint_as_float(a[j]) + V[ob + c + j] + res[j]; s1 += x; s2 = fmaf(x, x, s2); }
          }
          row_total(s1, s2);
          const float mean = s1 * (1.f / UD);
          const float rstd = 1.f / sqrtf(fmaxf(s2 * (1.f / UD) - mean * mean, 0.f) + 1e-5f);
          float cls[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) cls[q] = 0.f;
#pragma unroll 1
          for (int c = cb; c < cb + HC; c += 32) {
            uint32_t a[32];
            float res[32], y[32];
            tmem_ld32(tl + ocol + (uint32_t)c, a);
            residual(c, res);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = __uint_as_float(a[j]) + V[ob + c + j] + res[j];
              y[j] = (x - mean) * rstd * V[VEC_POST_W + c + j] + V[VEC_POST_B + c + j];
#pragma unroll
              for (int q = 0; q < 8; ++q) cls[q] = fmaf(y[j], V[VEC_CLS_W + q * UD + c + j], cls[q]);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 8) *a_chunk(&sm.C[0][0], r, c + j) = pack8(&y[j]);
          }
          if (layer == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) cls0[q] = cls[q];
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) cls1[q] = cls[q];
          }
        }
        TR();                                                             // 8 / 17: post_ln + classifier sums done
        to_mma();                                                         // -> G4 / G8 (gate)

        // gate: ReLU -> dropout -> Linear(128 -> 1)   (src/fusion.py:445-450)
        from_mma();
        TR();                                                             // 9 / 18: G4 / G8 arrived
        {
          float sc = 0.f;
#pragma unroll 1
          for (int c = half * (128 / NH); c < (half + 1) * (128 / NH); c += 32) {
            uint32_t a[32];
            tmem_ld32(tl + (uint32_t)c, a);
            tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 32; j4 += 4) {
              uint32_t wd[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
              if (P.ds.thr != 0u) {
                const Philox4 w = philox4x32_10_rk((uint32_t)((layer * 128 + c + j4) >> 2), gclip, pass, AVHOT_SITE_GATE, P.ds);
                wd[0] = w.x; wd[1] = w.y; wd[2] = w.z; wd[3] = w.w;
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float v = fmaxf(__uint_as_float(a[j4 + u]) + V[VEC_G0_B + c + j4 + u], 0.f);
                v = wd[u] >= P.ds.thr ? v * P.ds.scale : 0.f;
                sc = fmaf(v, V[VEC_G3_W + c + j4 + u], sc);
              }
            }
          }
          if (layer == 0) score0 = sc; else score1 = sc;                  // + gate bias after the halves are combined
        }
        TR();                                                             // 10 / 19: gate done
        if (layer == 0) { to_mma(); TR(); }                               // D_a drained -> G5 (Q_1 | K_1v); 11
      }

      // ---- modality softmax, finite guard, renormalise, logits (src/fusion.py:513-522) --------------------------------
      if (NH > 1) {
        // gate scores and classifier partial sums of the upper column half -> the lower half's thread of the same row,
        // through the C operand buffer (free: G8, its last reader, has completed; nobody writes it before the next
        // tile's G1 has been seen by every epilogue warp)
        float* mail = reinterpret_cast<float*>(&sm.C[0][0]) + r * (20 * (NH - 1));
        if (half != 0) {
          float* mine = mail + 20 * (half - 1);
          mine[0] = score0; mine[1] = score1;
#pragma unroll
          for (int q = 0; q < 8; ++q) { mine[2 + q] = cls0[q]; mine[10 + q] = cls1[q]; }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(32 * NH) : "memory");
        if (half == 0) {
#pragma unroll
          for (int h2 = 1; h2 < NH; ++h2) {
            const float* o = mail + 20 * (h2 - 1);
            score0 += o[0]; score1 += o[1];
#pragma unroll
            for (int q = 0; q < 8; ++q) { cls0[q] += o[2 + q]; cls1[q] += o[10 + q]; }
          }
        }
      }
      if (half == 0) {
        score0 += V[VEC_G3_B]; score1 += V[VEC_G3_B];
        float s0 = inv0 ? -1e4f : score0, s1 = inv1 ? -1e4f : score1;
        const float mx = fmaxf(s0, s1);
        const float e0 = __expf(s0 - mx), e1 = __expf(s1 - mx);
        float w0 = e0 / (e0 + e1), w1 = e1 / (e0 + e1);
        if (!isfinite(w0)) w0 = 0.f;
        if (!isfinite(w1)) w1 = 0.f;
        const float den = fmaxf(w0 + w1, 1e-8f);
        w0 /= den; w1 /= den;
        if (live) {
          float* out = P.logits + (size_t)row * P.C;
          for (int q = 0; q < P.C; ++q) out[q] = w0 * cls0[q] + w1 * cls1[q] + V[VEC_CLS_B + q];
        }
      }
      TR();                                                               // 20: logits written
      // the next tile's E0 overwrites X: every MMA that read X (G6) completed before G7's commit, which we have seen
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// out[(row0 + n) * 256 + k] = half(src_t[k * N + n]) for k < K, 0 for K <= k < 256   (src_t is [k][n])
__global__ void pack_rows_f16_kernel(const float* __restrict__ src_t, __half* __restrict__ out, int row0, int N, int K) {
  const int n = blockIdx.x;
  for (int k = threadIdx.x; k < UD; k += blockDim.x)
    out[(size_t)(row0 + n) * UD + k] = __float2half_rn(k < K ? src_t[(size_t)k * N + n] : 0.f);
}

}  // namespace tc

int fusion_tc_pack_rows(const float* src_t, void* packed, int row0, int N, int K, cudaStream_t s) {
  tc::pack_rows_f16_kernel<<<N, 128, 0, s>>>(src_t, reinterpret_cast<__half*>(packed), row0, N, K);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

size_t fusion_tc_packed_bytes() { return (size_t)tc::ROWS_ALL * tc::UD * sizeof(__half); }

int fusion_tc_forward(const void* packed_w, const FusTcParams& P, cudaStream_t s) {
  using namespace tc;
  CUtensorMap tmW;
  int rc = make_tensor_map_2d(&tmW, packed_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ROWS_ALL, UD, UD, 128, 64);
  if (rc != AVHOT_OK) return rc;
  const long long rows = (long long)P.B * P.T;
  AVHOT_REQUIRE(rows < 0x7fffffffLL, AVHOT_ERR_INVALID, "fusion_tc: B * T too large for one call");
  const int n_tiles = (int)((rows + UB - 1) / UB);
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  const int smem = (int)sizeof(FusTcSmem) + 1024;
  // AVHOT_FUSION_HALVES=1 selects the four-epilogue-warp form (thread = whole row) for A/B runs and tests
  static const int halves = [] {
    const char* e = getenv("AVHOT_FUSION_HALVES");
    return e && e[0] == '1' ? 1 : (e && e[0] == '4' ? 4 : 2);
  }();
  // debug: AVHOT_FUS_TRACE=<file> appends the clock64() stamps of one epilogue thread (fusion_tc_kernel, TR()) per launch
  static const char* trace_path = getenv("AVHOT_FUS_TRACE");
  FusTcParams PT = P;
  long long* d_trace = nullptr;
  if (trace_path && n_tiles >= 3 * grid) {
    AVHOT_CUDA(cudaMalloc(&d_trace, 64 * sizeof(long long)));
    AVHOT_CUDA(cudaMemsetAsync(d_trace, 0, 64 * sizeof(long long), s));
    PT.trace = d_trace;
  }
  struct TraceDump {
    long long* d; cudaStream_t s; const char* path;
    ~TraceDump() {
      if (!d) return;
      long long h[64];
      if (cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess)
        if (FILE* f = fopen(path, "a")) {
          for (int i = 1; i < 64 && h[i]; ++i) fprintf(f, "%d:%lld ", i, h[i] - h[i - 1]);
          fprintf(f, "| total %lld\n", h[20] - h[0]);
          fclose(f);
        }
      cudaFree(d);
    }
  } dump{d_trace, s, trace_path};
  ProfScope prof("fusion_tc", s);
  if (halves == 4) {
    AVHOT_CUDA(cudaFuncSetAttribute(fusion_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    fusion_tc_kernel<4><<<grid, 64 + 128 * 4, smem, s>>>(tmW, PT);
  } else if (halves == 2) {
    AVHOT_CUDA(cudaFuncSetAttribute(fusion_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    fusion_tc_kernel<2><<<grid, 64 + 128 * 2, smem, s>>>(tmW, PT);
  } else {
    AVHOT_CUDA(cudaFuncSetAttribute(fusion_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    fusion_tc_kernel<1><<<grid, 64 + 128, smem, s>>>(tmW, PT);
  }
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

}  // namespace avhot
