// fp32 LSTM recurrence for SMALL batches, weight-stationary over a thread-block cluster (AVHOT_PREC_FP32 path of a3,
// src/encoders.py:54-61, :148-154): the reference's configs[0] batch of 8, a default mc_forward call, the near-tie
// re-evaluation of a few hundred clips.
//
// lstm_recurrence_kernel (lstm.cu) streams all of W_hh (1 MB fp32) from L2 through shared memory EVERY step in 32 chunks
// with a block barrier each: 17 us per step whatever the batch, 10 ms for the 602 steps of two layers even for eight clips.
// Here a cluster of 8 CTAs keeps W_hh resident: CTA r owns hidden units [32 r, 32 r + 32) = 128 of the 1024 gate rows =
// 128 KB of shared memory, loaded once.  Every step each CTA computes its units' gates for ALL clips of the cluster
// (thread = (unit, clip group): the four gate weights of (k, unit) are one 128-bit shared load, h_{t-1}[k] of the group's
// clips a broadcast load), applies the activations, and writes its 32 units of h_t into the h buffer of every CTA of the
// cluster through distributed shared memory with st.async, which counts the bytes on the RECEIVER's mbarrier: a CTA starts
// step t + 1 when its own barrier has seen all 256 units of h_t - no cluster-wide barrier (a barrier.cluster round per step
// measured ~6 us, ten times the FMAs of eight clips).  h is double-buffered; a peer cannot overwrite the buffer a CTA is
// still reading, because producing h_{t+1} needs that CTA's own h_t first.  No weight traffic, no chunk barriers.
// The arithmetic and its order per (clip, unit, gate) - xp, then k = 0 .. 255 in sequence, expf / tanhf activations - are
// those of lstm_recurrence_kernel: results are bit-identical, so a clip's embedding does not depend on which kernel the
// batch size selected.
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace avhot {

constexpr int CSH = 256, CSU = 32, CSCL = 8;   // hidden size, units per CTA, cluster size

__device__ __forceinline__ float cs_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
// remote (or own) shared-memory store whose bytes complete the mbarrier at `bar` in the SAME CTA as `addr`
__device__ __forceinline__ void st_async_f32(uint32_t addr, float v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr),
               "r"(__float_as_uint(v)), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void st_async_f32x2(uint32_t addr, float a, float b, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(addr),
               "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void st_async_f32x4(uint32_t addr, float a, float b, float c, float d, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)), "r"(bar)
               : "memory");
}

// NB clips per cluster in NW groups (warps); thread = (unit u = tid & 31, clip group tid >> 5 of CT = NB / NW clips).
// Every warp reads the CTA's whole weight slice each step (4 shared-memory wavefronts per k), so large NB use few warps with
// many clips per thread (48 clips: 4 warps x 12 clips = 28 wavefronts per k; 8 warps x 6 = 44; measured equal, 15.6 us per
// step against 7 us of FMA issue - like the streaming kernel the FFMA stream runs at ~45 % of the pipe), small NB many
// warps (the FMA chain of one clip is latency-bound).
template <int NB, int NW>
__global__ void __launch_bounds__(32 * NW, 1)
lstm_rec_cs_kernel(const float* __restrict__ xp,      // (B, T, 4H)
                   const float* __restrict__ whh_t,   // [k][j][gate] = W_hh[gate*H + j][k]
                   float* __restrict__ h_seq,         // (B, T, H) or NULL
                   float* __restrict__ h_last,        // (B, H) or NULL
                   const int32_t* __restrict__ lengths, int B, int T) {
  constexpr int CT = NB / NW, CSTHREADS = 32 * NW;
  static_assert(NB % NW == 0, "clip groups");
  extern __shared__ __align__(16) unsigned char cs_smem_raw[];
  float4 (*ws)[CSU] = reinterpret_cast<float4 (*)[CSU]>(cs_smem_raw);                                    // [256 k][32 u] x 4 gates
  float (*hs)[CSH][NB] = reinterpret_cast<float (*)[CSH][NB]>(cs_smem_raw + sizeof(float4) * CSH * CSU);   // [2][k][clip]
  uint64_t* hbar = reinterpret_cast<uint64_t*>(cs_smem_raw + sizeof(float4) * CSH * CSU + sizeof(float) * 2 * CSH * NB);   // [2]
  constexpr uint32_t H_BYTES = CSH * NB * (uint32_t)sizeof(float);     // one step's h of the whole cluster batch
  const int tid = threadIdx.x, u = tid & 31, grp = tid >> 5;
  const uint32_t crank = tc::cluster_ctarank();
  const int cluster_id = blockIdx.x / CSCL;
  const int j = (int)crank * CSU + u;                 // this thread's hidden unit
  const int b0 = cluster_id * NB + grp * CT;          // its first clip

  // W_hh slice -> shared memory, once: for every k the 32 units x 4 gates of this CTA are 512 contiguous bytes
  for (int i = tid; i < CSH * CSU; i += CSTHREADS) {
    const int k = i / CSU, uu = i % CSU;
    ws[k][uu] = __ldg(reinterpret_cast<const float4*>(whh_t) + (size_t)k * CSH + crank * CSU + uu);
  }
  for (int i = tid; i < CSH * NB; i += CSTHREADS) (&hs[0][0][0])[i] = 0.f;     // h_{-1} = 0
  float c[CT];
#pragma unroll
  for (int b = 0; b < CT; ++b) c[b] = 0.f;
  // where this thread's h values go in every CTA of the cluster: hs[buf][j][grp * CT + b]
  uint32_t peer[CSCL], peer_bar[CSCL];
#pragma unroll
  for (int r = 0; r < CSCL; ++r) {
    peer[r] = tc::cluster_map_shared(&hs[0][j][grp * CT], (uint32_t)r);
    peer_bar[r] = tc::cluster_map_shared(&hbar[0], (uint32_t)r);
  }
  if (tid == 0) {
    tc::mbar_init(&hbar[0], 1);
    tc::mbar_init(&hbar[1], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  tc::cluster_sync_all();                             // every CTA's buffers are initialised before any remote write

  // xp of the next step is fetched while this step's activations and barrier run (its latency is a tenth of a step here)
  float xn[4][CT];
  auto fetch = [&](int t) {
#pragma unroll
    for (int b = 0; b < CT; ++b) {
      const float* x = xp + ((size_t)min(b0 + b, B - 1) * T + min(t, T - 1)) * (4 * CSH) + j;
#pragma unroll
      for (int g = 0; g < 4; ++g) xn[g][b] = __ldg(x + g * CSH);
    }
  };
  fetch(0);
  for (int t = 0; t < T; ++t) {
    const int cur = t & 1;
    float acc[4][CT];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int b = 0; b < CT; ++b) acc[g][b] = xn[g][b];
    fetch(t + 1);
    // arm the barrier that will collect h_t (buffer cur ^ 1), then wait for all of h_{t-1} (buffer cur; zeros at t = 0)
    if (tid == 0 && t + 1 < T) tc::mbar_expect_tx(&hbar[cur ^ 1], H_BYTES);
    if (t > 0) tc::mbar_wait_cluster(&hbar[cur], (uint32_t)((t - 1) >> 1) & 1u);
    // the next row's operands are fetched into a second register set before this row's FMAs (ptxas otherwise leaves each
    // shared load a few instructions in front of its use, which the few warps of a scheduler cannot cover)
    auto load_h = [&](int k, float (&hv)[CT]) {
      if (CT % 4 == 0) {
#pragma unroll
        for (int b4 = 0; b4 < CT; b4 += 4) {
          const float4 v = *reinterpret_cast<const float4*>(&hs[cur][k][grp * CT + b4]);
          hv[b4] = v.x; hv[b4 + 1] = v.y; hv[b4 + 2] = v.z; hv[b4 + 3] = v.w;
        }
      } else if (CT % 2 == 0) {
#pragma unroll
        for (int b2 = 0; b2 < CT; b2 += 2) {
          const float2 v = *reinterpret_cast<const float2*>(&hs[cur][k][grp * CT + b2]);
          hv[b2] = v.x; hv[b2 + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int b = 0; b < CT; ++b) hv[b] = hs[cur][k][grp * CT + b];
      }
    };
    float4 wq[2];
    float hq[2][CT];
    wq[0] = ws[0][u];
    load_h(0, hq[0]);
#pragma unroll 8
    for (int k = 0; k < CSH; ++k) {
      if (k + 1 < CSH) {
        wq[(k + 1) & 1] = ws[k + 1][u];
        load_h(k + 1, hq[(k + 1) & 1]);
      }
      const float4 w4 = wq[k & 1];
      const float wg[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int b = 0; b < CT; ++b) acc[g][b] = fmaf(wg[g], hq[k & 1][b], acc[g][b]);
    }
    float hn[CT];
#pragma unroll
    for (int b = 0; b < CT; ++b) {
      const float ig = cs_sigmoid(acc[0][b]), fg = cs_sigmoid(acc[1][b]);
      const float gg = tanhf(acc[2][b]), og = cs_sigmoid(acc[3][b]);
      c[b] = fmaf(fg, c[b], ig * gg);
      hn[b] = og * tanhf(c[b]);
    }
    // h_t of this thread's (unit, clips) -> buffer cur ^ 1 of all eight CTAs
    const uint32_t off = (uint32_t)((cur ^ 1) * CSH * NB) * (uint32_t)sizeof(float), boff = (uint32_t)(cur ^ 1) * 8u;
    if (t + 1 < T)                                    // (nobody reads the last step's h: every byte sent is waited for)
#pragma unroll
    for (int r = 0; r < CSCL; ++r) {
      const uint32_t dst = peer[r] + off, bar = peer_bar[r] + boff;
      if (CT % 4 == 0) {
#pragma unroll
        for (int b = 0; b < CT; b += 4) st_async_f32x4(dst + (uint32_t)b * 4u, hn[b], hn[b + 1], hn[b + 2], hn[b + 3], bar);
      } else if (CT % 2 == 0) {
#pragma unroll
        for (int b = 0; b < CT; b += 2) st_async_f32x2(dst + (uint32_t)b * 4u, hn[b], hn[b + 1], bar);
      } else {
#pragma unroll
        for (int b = 0; b < CT; ++b) st_async_f32(dst + (uint32_t)b * 4u, hn[b], bar);
      }
    }
    if (h_seq) {
#pragma unroll
      for (int b = 0; b < CT; ++b)
        if (b0 + b < B) h_seq[((size_t)(b0 + b) * T + t) * CSH + j] = hn[b];
    }
    if (h_last) {
#pragma unroll
      for (int b = 0; b < CT; ++b)
        if (b0 + b < B) {   // h_n = state after the clip's last valid step (packed sequences, src/encoders.py:139-147)
          const int last_t = (lengths ? min(max(__ldg(lengths + b0 + b), 1), T) : T) - 1;
          if (t == last_t) h_last[(size_t)(b0 + b) * CSH + j] = hn[b];
        }
    }
  }
  tc::cluster_sync_all();                             // nobody leaves while a peer may still read its barrier addresses
}

namespace {
template <int NB> int cs_smem_bytes() { return (int)(sizeof(float4) * CSH * CSU + sizeof(float) * 2 * CSH * NB) + 16; }

template <int NB, int NW> bool cs_try_launch(const float* xp, const float* whh_t, float* h_seq, float* h_last,
                                             const int32_t* lengths, int B, int T, cudaStream_t s, int* rc) {
  auto kern = lstm_rec_cs_kernel<NB, NW>;
  const int smem = cs_smem_bytes<NB>();
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(CSCL * ceil_div(B, NB)));
  cfg.blockDim = dim3(32 * NW);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CSCL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static const int max_clusters = [&] {               // co-resident clusters of this variant (one GPU per process)
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    return n;
  }();
  if (max_clusters < ceil_div(B, NB)) return false;   // a second wave would double the time: the streaming kernel then
  if (cudaLaunchKernelEx(&cfg, kern, xp, whh_t, h_seq, h_last, lengths, B, T) != cudaSuccess) {
    set_error("lstm_rec_cs: launch -> %s", cudaGetErrorString(cudaGetLastError()));
    *rc = AVHOT_ERR_CUDA;
    return true;
  }
  *rc = AVHOT_OK;
  return true;
}
}  // namespace

// Runs the recurrence of one layer with the cluster-resident kernel if the batch fits one wave of clusters at 8, 16, 32 or
// 48 clips per cluster; returns false (nothing launched) otherwise.  AVHOT_LSTM_CS=0 switches it off.
bool lstm_recurrence_cs(const float* xp, const float* whh_t, float* h_seq, float* h_last, const int32_t* lengths, int B,
                        int T, cudaStream_t s, int* rc) {
  static const bool off = [] { const char* e = getenv("AVHOT_LSTM_CS"); return e && atoi(e) == 0; }();
  if (off || B < 1) return false;
  if (cs_try_launch<8, 8>(xp, whh_t, h_seq, h_last, lengths, B, T, s, rc)) return true;
  if (cs_try_launch<16, 8>(xp, whh_t, h_seq, h_last, lengths, B, T, s, rc)) return true;
  if (cs_try_launch<32, 8>(xp, whh_t, h_seq, h_last, lengths, B, T, s, rc)) return true;
  if (cs_try_launch<48, 4>(xp, whh_t, h_seq, h_last, lengths, B, T, s, rc)) return true;
  return false;
}

}  // namespace avhot
