// One LSTM layer as ONE kernel on the tcgen05 tensor cores: input projection and recurrence in the same MMA chain
// (AVHOT_PREC_TC path of a3, src/encoders.py:54-61, :148-154; SURVEY.md section 7 "hard parts").
//
// lstm_tc.cu needs xp = x W_ih^T + b for all steps in HBM first (gemm_ws.cu): 12 GB written and read back per layer at
// 18 944 clips, 26x the algorithmic traffic of the stage, and a second 3 ms kernel.  Here every step computes
//     gates[256 clips, 1024] = [x_t | h_{t-1}] x [W_ih | W_hh]^T            (K = KX * 64 + 256, one accumulator chain)
// per CTA pair (cta_group::2, M = 256 = 128 clips per CTA), in 8 chunks of 32 hidden units (N = 128 accumulator
// columns, two TMEM buffers).  What made it fit (round 2 wrote it off: 128 KB of h ping-pong + 64 KB of x_t + the
// weight ring > 227 KB of shared memory, TMEM full):
//   * h lives in TENSOR MEMORY, not shared memory: the epilogue threads write h_t as packed fp16 with tcgen05.st into
//     128 columns of their own lanes, and the recurrent half of the chain is tcgen05.mma with the A operand FROM TMEM
//     ([a_tmem] form, K = 16 = 8 columns per instruction); two buffers (step parity) = 256 columns;
//   * the cell state moves from TMEM to REGISTERS: thread (clip, q) owns 8 units of each of the 8 chunks = 64 fp32
//     registers, statically indexed because the chunk loop is unrolled; the xp prefetch registers are gone;
//   * shared memory then holds only the x_t tiles (a ring of NXS 16 KB k-block slots, filled by TMA from the time-major
//     fp16 operand: the front-end's [hi | lo | hi] split for layer 0, K = 320; the previous layer's hidden sequence
//     above it, K = 256), the weight ring (NWS stages of G k-blocks, this CTA's half of the gate rows, streamed from
//     L2 every step as before) and, below the top layer, two 16 KB staging tiles from which the layer's own fp16
//     hidden sequence leaves through TMA stores (the next layer's x operand).
// The x half of chunk 0 of step t+1 does not depend on h_t, so the tensor pipe works through it while the epilogue
// of step t's last chunk finishes: the step boundary costs no bubble.  TMEM: [0,256) accumulators, [256,512) h.
//   warp 0      TMA producer: weight stages (W_ih chunk c, then W_hh chunk c), both CTAs, completing the leader's barriers
//   warp 1      leader CTA: one elected lane issues the MMAs of a chunk (4 KX SS on x_t, 16 TS on h_{t-1})
//   warps 4-19  epilogue: thread = (clip = TMEM lane, q): tcgen05.ld 4 gates x 8 units, + bias, activations, c in
//               registers, h_t -> tcgen05.st, arrive on h_ready[k-block] of the leader
//   warp 2      TMA producer of the x_t k-blocks (runs ahead by the free slots of the ring)
//   warp 3      TMA stores of the layer's hidden sequence (HSEQ; costs 0.33 ms of layer 0's 3.6 - writing the 16 bytes of a
//               thread straight to global memory instead was measured slower still: 7.66 against 7.29 ms for both layers)
// Gate columns of a chunk are permuted so that a thread's (4 gates x 4 units) are 16 consecutive TMEM columns:
//   packed row c*128 + q*32 + hf*16 + g*4 + i  =  gate g, unit c*32 + q*8 + hf*4 + i     (pack_fused_w_kernel)
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace avhot {
namespace tc {

constexpr int FB = 128, FH = 256, FCH = 32, FNC = FH / FCH, FN = 4 * FCH, FHKB = FH / 64;
constexpr int FTHREADS = 128 + 512;            // warpgroup 0: producers + MMA (setmaxnreg 56), warpgroups 1-4: epilogue (104): 128 x 56 + 512 x 104 <= 640 x 96, the launch allocation
constexpr uint32_t FCOL_H = 256;                 // first TMEM column of the hidden state (2 x 128 columns)

template <int NXS, int NWS, int G, int NHS> struct FusedSmem {   // G = k-blocks a weight stage can hold
  __half xA[NXS][FB * 64];          // x_t k-blocks, A operand: 128 rows x 128 B, 128-byte swizzle
  __half wB[NWS][G][64 * 64];       // weight stages: G k-blocks of (this CTA's 64 gate rows x 64 k), B operand
  __half hS[NHS ? NHS : 1][NHS ? FB * 64 : 8];   // HSEQ: k-blocks of h_t on their way out (TMA store source)
  uint64_t w_full[NWS], w_empty[NWS], acc_full[2], acc_empty[2], h_ready[FHKB], x_full[NXS], x_empty[NXS];
  uint64_t hs_done[2], hs_free[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait_f() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T, CTA pair: each CTA's 128 rows of A are 128 lanes x 8 columns (K = 16 fp16, two per
// 32-bit column, row = lane) of its own tensor memory at the same address
__device__ __forceinline__ void umma_f16_pair_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                                       bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// tanh.approx.f32: one MUFU, max relative error 2^-11.  (ex2 + rcp activations, two MUFU and ~1e-7, were measured: the
// audio path's share of the logit error stayed at 6e-6 rms / 4.8e-5 max of the logit scale - it is set by the fp16
// hidden state and weights - and the kernel went from 7.3 to 9.0 ms per step.)
__device__ __forceinline__ float tanh_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid of a pre-activation that arrives halved (xh = x / 2)
__device__ __forceinline__ float sigmoid_half(float xh) { return fmaf(0.5f, tanh_mufu(xh), 0.5f); }
__device__ __forceinline__ void stg256f(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// KX: 64-column k-blocks of the x operand (5 = the [hi | lo | hi] split of 104 features, 3 = of the 64 log-mel features
// alone with the MFCC's DCT folded into the weights, 4 = a 256-wide hidden sequence).  G: k-blocks per weight stage (divides KX and 4).  HSEQ: the layer's fp16 hidden sequence is stored
// (time-major (T, Bp, pitch), rows of padding clips included) and h_last is not.  The sigmoid gates' weights and
// biases are packed halved (tc_gate_scale in lstm.cu): sigmoid(x) = 0.5 + 0.5 tanh(x / 2), one MUFU per activation.
template <int KX, int NXS, int NWS, int GX, int GH, bool HSEQ>
__global__ void __launch_bounds__(FTHREADS, 1)
lstm_fused_pair_kernel(const __grid_constant__ CUtensorMap tmWi, const __grid_constant__ CUtensorMap tmWh,
                       const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH,
                       const float* __restrict__ bias_p, float* __restrict__ h_last,
                       const int32_t* __restrict__ lengths, int B, int T, int n_ct_arg) {
  static_assert(KX % GX == 0 && FHKB % GH == 0, "a weight stage holds GX / GH k-blocks of the x / h half of a chunk");
  constexpr int G = GX > GH ? GX : GH;
  static_assert(NXS > KX, "x ring: a whole step plus at least one slot to load ahead into");
  using Smem = FusedSmem<NXS, NWS, (GX > GH ? GX : GH), HSEQ ? 2 : 0>;
  const int ctile = min((int)blockIdx.x, n_ct_arg - 1);
  const bool writer = (int)blockIdx.x < n_ct_arg;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  extern __shared__ unsigned char fused_smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(fused_smem_raw + ((1024u - (smem_u32(fused_smem_raw) & 1023u)) & 1023u));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmWi);
    tma_prefetch_desc(&tmWh);
    for (int s = 0; s < NWS; ++s) { mbar_init(&sm.w_full[s], 1); mbar_init(&sm.w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sm.acc_full[b], 1); mbar_init(&sm.acc_empty[b], 16 * 2); }
    for (int k = 0; k < FHKB; ++k) mbar_init(&sm.h_ready[k], 32 * 2);   // 16 epilogue warps x 2 chunks x 2 CTAs
    for (int s = 0; s < NXS; ++s) { mbar_init(&sm.x_full[s], 1); mbar_init(&sm.x_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sm.hs_done[b], 32); mbar_init(&sm.hs_free[b], 1); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_pair(&sm.tmem_base, 512); tmem_relinquish_pair(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                          // both CTAs' barriers are initialised before any remote arrival
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  constexpr uint32_t W_KB_BYTES = 2u * 64u * 64u * (uint32_t)sizeof(__half);   // one k-block, both CTAs' halves
  constexpr uint32_t X_SLOT_BYTES = 2u * FB * 64u * (uint32_t)sizeof(__half);
  constexpr int XST = KX / GX, HST = FHKB / GH;  // weight stages of the x half / the h half of a chunk

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ---- weight producer (both CTAs): per (t, c) the x half (W_ih rows of chunk c), then the h half (W_hh; none at t = 0)
      if (lane == 0) {
        const uint32_t full0 = cluster_map_shared(&sm.w_full[0], 0);      // the leader's w_full[] barriers
        int s = 0;
        uint32_t ph = 1;                        // parity to wait for on w_empty[s]
        for (int t = 0; t < T; ++t)
          for (int c = 0; c < FNC; ++c) {
            const int nst = t == 0 ? XST : XST + HST;
            for (int st = 0; st < nst; ++st) {
              mbar_wait(&sm.w_empty[s], ph);
              const bool xh = st < XST;
              if (leader) mbar_expect_tx(&sm.w_full[s], (xh ? GX : GH) * W_KB_BYTES);
              const CUtensorMap* tm = xh ? &tmWi : &tmWh;
              const int kb0 = xh ? st * GX : (st - XST) * GH;
#pragma unroll
              for (int j = 0; j < G; ++j)
                if (j < (xh ? GX : GH))
                  tma_load_2d_pair(sm.wB[s][j], tm, full0 + (uint32_t)s * 8u, (kb0 + j) * 64, c * FN + (int)crank * 64);
              if (++s == NWS) { s = 0; ph ^= 1u; }
            }
          }
      }
    } else if (warp == 2) {
      // ---- x producer (both CTAs): k-block kb of step t -> slot (t * KX + kb) % NXS, as far ahead as slots are free
      if (lane == 0) {
        tma_prefetch_desc(&tmX);
        const uint32_t xfull0 = cluster_map_shared(&sm.x_full[0], 0);
        int slot = 0;
        uint32_t ph = 1;
        for (int t = 0; t < T; ++t)
          for (int kb = 0; kb < KX; ++kb) {
            mbar_wait(&sm.x_empty[slot], ph);
            if (leader) mbar_expect_tx(&sm.x_full[slot], X_SLOT_BYTES);
            tma_load_2d_pair(sm.xA[slot], &tmX, xfull0 + (uint32_t)slot * 8u, kb * 64, (t * n_ct_arg + ctile) * FB);
            if (++slot == NXS) { slot = 0; ph ^= 1u; }
          }
      }
    } else if (warp == 1) {
      if (leader) {                             // the whole warp, converged; one elected lane issues
        constexpr uint32_t idesc = umma_idesc(FMT_F16, 2 * FB, FN);
        int s = 0, xbase = 0;                   // weight stage; slot of x_t's k-block 0
        uint32_t wph = 0, xph = 0;              // parities of w_full[s] / of x_full[xbase]
        for (int t = 0; t < T; ++t) {
          const uint32_t hsrc = tmem + FCOL_H + (uint32_t)(t & 1) * 128u;     // h_{t-1}
#pragma unroll 1
          for (int c = 0; c < FNC; ++c) {
            const int b = c & 1;
            // buffer b is used FNC / 2 = 4 times per step: the parity of its use count at chunk c is (c >> 1) & 1
            mbar_wait_cluster(&sm.acc_empty[b], (uint32_t)((c >> 1) & 1) ^ 1u);
            tc_fence_after();
            const uint32_t acc = tmem + (uint32_t)(b * FN);
#pragma unroll
            for (int st = 0; st < XST; ++st) {  // x half: A = x_t (shared memory), B = W_ih rows of chunk c
              mbar_wait(&sm.w_full[s], wph);
              tc_fence_after();
#pragma unroll
              for (int j = 0; j < GX; ++j) {
                const int kb = st * GX + j;
                int slot = xbase + kb;
                uint32_t ph = xph;
                if (slot >= NXS) { slot -= NXS; ph ^= 1u; }
                if (c == 0) {
                  mbar_wait(&sm.x_full[slot], ph);
                  tc_fence_after();
                }
                const uint64_t adesc = umma_desc_sw128(smem_u32(sm.xA[slot]));
                const uint64_t bdesc = umma_desc_sw128(smem_u32(sm.wB[s][j]));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_pair_elect(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                if (c == FNC - 1) umma_commit_pair_mc_elect(&sm.x_empty[slot], 3);   // last reader of x_t's k-block
              }
              umma_commit_pair_mc_elect(&sm.w_empty[s], 3);
              if (++s == NWS) { s = 0; wph ^= 1u; }
            }
            if (t > 0) {                        // h half: A = h_{t-1} (tensor memory), B = W_hh rows of chunk c
#pragma unroll
              for (int st = 0; st < HST; ++st) {
                mbar_wait(&sm.w_full[s], wph);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < GH; ++j) {
                  const int kb = st * GH + j;
                  if (c == 0) {
                    mbar_wait_cluster(&sm.h_ready[kb], (uint32_t)(t - 1) & 1);
                    tc_fence_after();
                  }
                  const uint64_t bdesc = umma_desc_sw128(smem_u32(sm.wB[s][j]));
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16_pair_ts_elect(acc, hsrc + (uint32_t)(kb * 32 + k * 8), bdesc + 2 * k, idesc, true);
                }
                umma_commit_pair_mc_elect(&sm.w_empty[s], 3);
                if (++s == NWS) { s = 0; wph ^= 1u; }
              }
            }
            umma_commit_pair_mc_elect(&sm.acc_full[b], 3);
          }
          xbase += KX;
          if (xbase >= NXS) { xbase -= NXS; xph ^= 1u; }
        }
      }
    } else if (HSEQ) {
      // ---- store warp: k-block kb of h_t (128 clips x 64 units, swizzled like a TMA box) -> the time-major fp16 sequence
      if (lane == 0) tma_prefetch_desc(&tmH);
      for (int t = 0; t < T; ++t)
        for (int kb = 0; kb < FHKB; ++kb) {
          const int buf = kb & 1;
          mbar_wait(&sm.hs_done[buf], (uint32_t)(kb >> 1) & 1);            // two uses of a buffer per step
          if (lane == 0) {
            if (writer) {
              tma_store_2d(&tmH, sm.hS[buf], kb * 64, (t * n_ct_arg + ctile) * FB);
              tma_store_commit();
              tma_store_wait_read0();
            }
            mbar_arrive(&sm.hs_free[buf]);
          }
          __syncwarp();
        }
      if (lane == 0) tma_store_wait0();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ---- 16 epilogue warps = (TMEM lane quarter) x (quarter q of a chunk); thread (clip r, q) owns units
    // 32 c + 8 q .. + 7 of every chunk c and keeps their cell state in registers for the whole sequence
    const int q = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // clip row = TMEM lane
    const int clip = ctile * FB + r;
    const bool live = writer && clip < B;
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
    const int last_t = (live && lengths ? min(max(__ldg(lengths + clip), 1), T) : T) - 1;
    const uint32_t acc_empty_addr0 = cluster_map_shared(&sm.acc_empty[0], 0);
    const uint32_t h_ready_addr0 = cluster_map_shared(&sm.h_ready[0], 0);
    const float4* const bias4 = reinterpret_cast<const float4*>(bias_p) + q * 8;
    unsigned char* const hs_row = reinterpret_cast<unsigned char*>(&sm.hS[0][0]) + (size_t)r * 128;
    float cs[FNC * 8];
#pragma unroll
    for (int i = 0; i < FNC * 8; ++i) cs[i] = 0.f;
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const uint32_t hdst = tlane + FCOL_H + (uint32_t)((t + 1) & 1) * 128u + (uint32_t)(q * 4);
      const bool emit = !HSEQ && live && h_last && t == last_t;
#pragma unroll
      for (int c = 0; c < FNC; ++c) {
        const int b = c & 1;
        mbar_wait(&sm.acc_full[b], (uint32_t)((c >> 1) & 1));
        tc_fence_after();
        uint32_t a[2][16];
        tmem_ld16(tlane + (uint32_t)(b * FN + q * 32), a[0]);
        tmem_ld16(tlane + (uint32_t)(b * FN + q * 32 + 16), a[1]);
        tmem_ld_wait();
        // the accumulator is in registers: hand the TMEM buffer back to the MMA warp before the activation math
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_addr0 + (uint32_t)b * 8u);
        float hn[8];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float4 bg[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) bg[g] = __ldg(bias4 + c * 32 + hf * 4 + g);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float bi = i == 0 ? bg[0].x : i == 1 ? bg[0].y : i == 2 ? bg[0].z : bg[0].w;
            const float bf = i == 0 ? bg[1].x : i == 1 ? bg[1].y : i == 2 ? bg[1].z : bg[1].w;
            const float bgg = i == 0 ? bg[2].x : i == 1 ? bg[2].y : i == 2 ? bg[2].z : bg[2].w;
            const float bo = i == 0 ? bg[3].x : i == 1 ? bg[3].y : i == 2 ? bg[3].z : bg[3].w;
            const float ig = sigmoid_half(__uint_as_float(a[hf][i]) + bi);
            const float fg = sigmoid_half(__uint_as_float(a[hf][4 + i]) + bf);
            const float gg = tanh_mufu(__uint_as_float(a[hf][8 + i]) + bgg);
            const float og = sigmoid_half(__uint_as_float(a[hf][12 + i]) + bo);
            const float cn = fmaf(fg, cs[c * 8 + hf * 4 + i], ig * gg);
            cs[c * 8 + hf * 4 + i] = cn;
            hn[hf * 4 + i] = og * tanh_mufu(cn);
          }
        }
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half2 v = __floats2half2_rn(hn[2 * j], hn[2 * j + 1]);
          pk[j] = *reinterpret_cast<const uint32_t*>(&v);
        }
        tmem_st4(hdst + (uint32_t)(c * 16), pk);             // units 32 c + 8 q .. + 7 = columns 16 c + 4 q .. + 3
        if (emit) stg256f(h_last + (size_t)clip * FH + c * FCH + q * 8, hn);
        if (HSEQ) {
          // the same 16 bytes into the staging tile of k-block c / 2 (buffer (c / 2) & 1), XOR-swizzled by the row
          const int buf = (c >> 1) & 1, use = c >> 2;          // second half of the step = second use of the buffer
          if ((c & 1) == 0) {
            if (use == 0) { if (t > 0) mbar_wait(&sm.hs_free[buf], 1u); }   // freed after use 1 of step t - 1
            else mbar_wait(&sm.hs_free[buf], 0u);                           // freed after use 0 of this step
          }
          const int chunk16 = (c & 1) * 4 + q;                 // 16-byte chunk of the 128-byte row: units (c & 1) 32 + 8 q
          *reinterpret_cast<uint4*>(hs_row + (size_t)buf * (FB * 128) + ((chunk16 ^ (r & 7)) << 4)) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
          fence_proxy_async_smem();
        }
        tmem_st_wait_f();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(h_ready_addr0 + (uint32_t)(c >> 1) * 8u);
          if (HSEQ) mbar_arrive(&sm.hs_done[(c >> 1) & 1]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                          // no CTA leaves while its peer can still signal its barriers
  if (warp == 1) tmem_dealloc_pair(tmem, 512);
}

}  // namespace tc

namespace {
// KX = 5 with a hidden sequence out (layer 0 on the front-end's split operand), KX = 4 top layer, KX = 4 middle layer
template <int KX, bool HSEQ> struct FusedCfg;
template <> struct FusedCfg<5, true> { static constexpr int NXS = 7, NWS = 10, GX = 1, GH = 1; };
template <> struct FusedCfg<5, false> { static constexpr int NXS = 8, NWS = 12, GX = 1, GH = 1; };
// log-mel split only (DCT folded): whole halves of a chunk per stage (the MMA thread waits and commits twice per chunk;
// per-k-block stages cost it ~45 cycles of tensor-pipe bubble each: 3.8 ms against 3.4 for this layer)
template <> struct FusedCfg<3, true> { static constexpr int NXS = 6, NWS = 3, GX = 3, GH = 4; };
template <> struct FusedCfg<3, false> { static constexpr int NXS = 6, NWS = 4, GX = 3, GH = 4; };
template <> struct FusedCfg<4, false> { static constexpr int NXS = 8, NWS = 3, GX = 4, GH = 4; };
template <> struct FusedCfg<4, true> { static constexpr int NXS = 8, NWS = 4, GX = 2, GH = 2; };
template <int KX, bool HSEQ> struct FusedTypes {
  using C = FusedCfg<KX, HSEQ>;
  static constexpr int G = C::GX > C::GH ? C::GX : C::GH;
  using Smem = tc::FusedSmem<C::NXS, C::NWS, G, HSEQ ? 2 : 0>;
  static constexpr int smem_bytes = (int)sizeof(Smem) + 1024;
  static_assert(smem_bytes <= 227 * 1024, "one CTA per SM");
  static auto kernel() { return tc::lstm_fused_pair_kernel<KX, C::NXS, C::NWS, C::GX, C::GH, HSEQ>; }
};

void fused_cluster_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int n_ct, int smem, cudaStream_t s) {
  cfg = {};
  cfg.gridDim = dim3((unsigned)(ceil_div(n_ct, 2) * 2));
  cfg.blockDim = dim3(tc::FTHREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = s;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

// can a CTA pair of this variant be resident on this device at all (one process drives one GPU: queried once)
template <int KX, bool HSEQ> bool fused_fits(int n_ct) {
  static const int max_clusters = [] {
    auto kern = FusedTypes<KX, HSEQ>::kernel();
    const int smem = FusedTypes<KX, HSEQ>::smem_bytes;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    fused_cluster_cfg(cfg, attr, 2 * 1024, smem, nullptr);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
  }();
  return max_clusters >= ceil_div(n_ct, 2);
}

template <int KX, bool HSEQ>
int fused_launch(const void* w_ih_p16, const void* w_hh_p16, const float* bias_p, const void* x_f16, int x_pitch,
                 void* h_seq_f16, int h_pitch, float* h_last, const int32_t* lengths, int B, int T, cudaStream_t s) {
  using namespace tc;
  const int n_ct = ceil_div(B, FB);
  const int smem = FusedTypes<KX, HSEQ>::smem_bytes;
  CUtensorMap tmWi, tmWh, tmX, tmH;
  int rc = make_tensor_map_2d(&tmWi, w_ih_p16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 4 * FH, KX * 64, KX * 64, 64, 64);
  if (rc == AVHOT_OK) rc = make_tensor_map_2d(&tmWh, w_hh_p16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 4 * FH, FH, FH, 64, 64);
  if (rc == AVHOT_OK)
    rc = make_tensor_map_2d(&tmX, x_f16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)T * n_ct * FB, KX * 64,
                            (uint64_t)x_pitch, FB, 64);
  tmH = tmX;                                    // unused without an output sequence
  if (rc == AVHOT_OK && HSEQ)
    rc = make_tensor_map_2d(&tmH, h_seq_f16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)T * n_ct * FB, FH,
                            (uint64_t)h_pitch, FB, 64);
  if (rc != AVHOT_OK) return rc;
  auto kern = FusedTypes<KX, HSEQ>::kernel();
  AVHOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fused_cluster_cfg(cfg, attr, n_ct, smem, s);
  ProfScope prof("lstm_layer_fused_tc", s);
  AVHOT_CUDA(cudaLaunchKernelEx(&cfg, kern, tmWi, tmWh, tmX, tmH, bias_p, h_last, lengths, B, T, n_ct));
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}
}  // namespace

// Does this layer run as the fused kernel?  Whenever the x operand has 5 (the split of <= 104 features), 3 (log-mel
// split, DCT folded) or 4 (a hidden sequence) k-blocks - at EVERY batch size, so that a clip's embedding does not depend
// on how the batch was sharded (the two-kernel path rounds xp to fp16: same tolerance, different bits).  A batch of
// one tile runs as a pair whose second CTA duplicates the tile and writes nothing; more pairs than fit on the GPU run in
// waves.  Small batches pay for it: a fused step is a dependent chain of 8 x (4 KX + 16) MMAs whatever the batch
// (3.4 / 3.7 / 4.4 ms per layer for KX = 3 / 4 / 5 at T = 301), while the two-kernel path's input GEMM spreads over all
// SMs (B = 512: 7.2 against 6.2 ms for the whole forward).  AVHOT_LSTM_FUSED=0: the two-kernel path everywhere.
bool lstm_fused_tc_supported(int x_cols, bool hseq, int B) {
  static const int forced = [] { const char* e = getenv("AVHOT_LSTM_FUSED"); return e ? atoi(e) : -1; }();
  if (forced == 0 || B < 1) return false;
  if (x_cols == 192) return hseq ? fused_fits<3, true>(1) : fused_fits<3, false>(1);
  if (x_cols == 320) return hseq ? fused_fits<5, true>(1) : fused_fits<5, false>(1);
  if (x_cols == 256) return hseq ? fused_fits<4, true>(1) : fused_fits<4, false>(1);
  return false;
}

// x_f16: (T, Bp, x_pitch) fp16 time-major operand (Bp = B rounded up to 128) of x_cols columns; weights / bias in the
// packed row order (fused_src_row).  h_seq_f16 != NULL: the layer's hidden sequence (T, Bp, h_pitch) fp16 is stored;
// else h_last (B, 256) fp32 = the hidden state at step lengths[b] - 1.
int lstm_layer_fused_tc(const void* w_ih_p16, const void* w_hh_p16, const float* bias_p, const void* x_f16, int x_cols,
                        int x_pitch, void* h_seq_f16, int h_pitch, float* h_last, const int32_t* lengths, int B, int T,
                        cudaStream_t s) {
  AVHOT_REQUIRE(x_pitch >= x_cols && (x_pitch & 7) == 0, AVHOT_ERR_INVALID, "lstm_layer_fused_tc: bad x pitch %d", x_pitch);
  AVHOT_REQUIRE(!h_seq_f16 || (h_pitch >= tc::FH && (h_pitch & 7) == 0), AVHOT_ERR_INVALID,
                "lstm_layer_fused_tc: bad h_seq pitch %d", h_pitch);
  AVHOT_REQUIRE(h_seq_f16 || h_last, AVHOT_ERR_INVALID, "lstm_layer_fused_tc: no output");
#define AVHOT_FUSED_CASE(KX_, HS_) \
  return fused_launch<KX_, HS_>(w_ih_p16, w_hh_p16, bias_p, x_f16, x_pitch, h_seq_f16, h_pitch, h_last, lengths, B, T, s)
  if (x_cols == 192 && h_seq_f16) AVHOT_FUSED_CASE(3, true);
  if (x_cols == 192) AVHOT_FUSED_CASE(3, false);
  if (x_cols == 320 && h_seq_f16) AVHOT_FUSED_CASE(5, true);
  if (x_cols == 320) AVHOT_FUSED_CASE(5, false);
  if (x_cols == 256 && h_seq_f16) AVHOT_FUSED_CASE(4, true);
  if (x_cols == 256) AVHOT_FUSED_CASE(4, false);
#undef AVHOT_FUSED_CASE
  set_error("lstm_layer_fused_tc: x operand of %d columns", x_cols);
  return AVHOT_ERR_UNSUPPORTED;
}

}  // namespace avhot
