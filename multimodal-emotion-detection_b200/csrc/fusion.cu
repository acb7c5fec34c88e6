// a6 + a7 (+ the a5 instances inside a6): HybridFusion.forward (src/fusion.py:455-530) with its two
// CrossModalAttention layers (src/attention.py:80-141, Tq = 1, Tk = M = 2), and the live baseline head
// (src/train.py:81-85, :153-164), for R = B*T rows (clip, stochastic pass).
//
// One CTA = 16 rows, 256 threads, thread n owns output column n of every D = 256 wide linear layer, so
// a layer is "for k: w = Wt[k][n] (coalesced, L1/L2 resident); acc[row] += w * x[row][k] (shared-memory
// broadcast)".  Everything between the layers stays on chip:
//   proj (dim_m -> 256) -> pre_ln -> [per layer i: q/k/v projections (the three input LayerNorms share
//   their statistics with each other and across layers, so x^ = (Z - mean)/std is formed once and the
//   LayerNorm affines are folded into the projection weights at create time) -> 4-head scores over the 2
//   modality tokens -> mask(-1e4) -> softmax -> all-masked rows to 0 -> Philox dropout -> PV in registers
//   -> out_proj -> + residual -> post_ln] -> gate MLP (Philox dropout) -> masked softmax over modalities
//   -> finite guard + renormalise -> weighted sum -> classifier.
#include "common.cuh"
#include "fusion_tc.cuh"

struct avhot_fusion {
  int kind;             // 0 hybrid, 1 concat head
  int M, D, heads, C, G;
  int dim[AVHOT_MAX_MODALITIES];
  // hybrid
  float* proj_wt[AVHOT_MAX_MODALITIES];   // [k][n]
  float* proj_b[AVHOT_MAX_MODALITIES];
  float *pre_w, *pre_b, *post_w, *post_b;
  float* qkv_wt[AVHOT_MAX_MODALITIES][3];  // LN-folded, [k][n]
  float* qkv_b[AVHOT_MAX_MODALITIES][3];   // LN-folded
  float* o_wt[AVHOT_MAX_MODALITIES];
  float* o_b[AVHOT_MAX_MODALITIES];
  float *g0_wt, *g0_b, *g3_w, *g3_b;       // g0_wt [k][g]
  float *cls_w, *cls_b;                    // (C, D)
  // concat head
  int head_in, head_hidden;
  float *h0_wt, *h0_b;                     // [k][n]
  void* tc_w;                              // packed fp16 weights of the tensor-core per-pass kernel (fusion_tc.cu)
  float* tc_vec;                           // its small fp32 vectors
  float* arena;
};

namespace avhot {

constexpr int FD = 256, FHEADS = 4, FDH = 64, FG = 128, FM = 2, FRT = 16, FTHREADS = 256, FCMAX = 16;

// Wt[k][n] = W[n][k] * ln_w[k];  bf[n] = b[n] + sum_k W[n][k] * ln_b[k]   (ln_w == NULL: plain transpose)
__global__ void fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ b,
                               const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                               float* __restrict__ Wt, float* __restrict__ bf, int N, int K) {
  const int n = blockIdx.x;
  double acc = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float w = W[(size_t)n * K + k];
    Wt[(size_t)k * N + n] = ln_w ? w * ln_w[k] : w;
    if (ln_b) acc += (double)w * (double)ln_b[k];
  }
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    bf[n] = (float)((double)b[n] + s);
  }
}

struct FusionParams {
  const float* z[FM];
  int rpc[FM];            // rows per clip of z[m]: 1 (pass-invariant) or T
  int dim[FM];
  const float* mask;      // (B, M) fp32, 1 = available, or NULL
  float *logits, *fusion_weights, *attn, *h_att;
  int B, T, C;
  const float *proj_wt[FM], *proj_b[FM];
  const float *pre_w, *pre_b, *post_w, *post_b;
  const float *qkv_wt[FM][3], *qkv_b[FM][3], *o_wt[FM], *o_b[FM];
  const float *g0_wt, *g0_b, *g3_w, *g3_b, *cls_w, *cls_b;
  DropSpec ds;
};

struct FusionSmem {
  float xh[FRT][FM][FD];      // x^ = (Z - mean) / std of the pre_ln output Z
  float hatt[FRT][FM][FD];    // H_att
  float buf[FRT][FD];         // ctx, then fused
  float zmean[FRT][FM], zstd[FRT][FM];
  float part[8][FRT][FM];     // per-warp partial reductions
  float attn[FRT][FHEADS][FM];
  float fw[FRT][FM];
};

// LayerNorm of one 256-vector held as 8 values per lane (columns lane + 32 q); returns mean / rstd
__device__ __forceinline__ void warp_ln_stats(const float (&v)[8], float& mean, float& var) {
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += v[q];
  mean = warp_sum(s) * (1.f / FD);
  float sq = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) { const float d = v[q] - mean; sq = fmaf(d, d, sq); }
  var = warp_sum(sq) * (1.f / FD);
}

__global__ void __launch_bounds__(FTHREADS, 2) hybrid_fusion_kernel(const FusionParams P) {
  extern __shared__ __align__(16) unsigned char fusion_smem_raw[];
  FusionSmem& sm = *reinterpret_cast<FusionSmem*>(fusion_smem_raw);
  const int n = threadIdx.x, lane = n & 31, warp = n >> 5;
  const long long n_rows = (long long)P.B * P.T;
  const long long row0 = (long long)blockIdx.x * FRT;

  // ---- phase 1: Z = pre_ln(proj_m(z_m)); keep x^ and the statistics of Z ------------------------------
#pragma unroll
  for (int m = 0; m < FM; ++m) {
    float acc[FRT];
#pragma unroll
    for (int r = 0; r < FRT; ++r) acc[r] = 0.f;
    const float* src[FRT];
#pragma unroll
    for (int r = 0; r < FRT; ++r) {
      long long row = row0 + r;
      if (row >= n_rows) row = n_rows - 1;
      const long long b = row / P.T, t = row % P.T;
      src[r] = P.z[m] + (size_t)(b * P.rpc[m] + (P.rpc[m] == 1 ? 0 : t)) * P.dim[m];
    }
    const float* wt = P.proj_wt[m] + n;
    for (int k = 0; k < P.dim[m]; k += 4) {
      const float w0 = __ldg(wt + (size_t)k * FD), w1 = __ldg(wt + (size_t)(k + 1) * FD);
      const float w2 = __ldg(wt + (size_t)(k + 2) * FD), w3 = __ldg(wt + (size_t)(k + 3) * FD);
#pragma unroll
      for (int r = 0; r < FRT; ++r) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(src[r] + k));
        acc[r] = fmaf(w0, x.x, acc[r]); acc[r] = fmaf(w1, x.y, acc[r]);
        acc[r] = fmaf(w2, x.z, acc[r]); acc[r] = fmaf(w3, x.w, acc[r]);
      }
    }
    const float pb = __ldg(P.proj_b[m] + n);
#pragma unroll
    for (int r = 0; r < FRT; ++r) sm.xh[r][m][n] = acc[r] + pb;
  }
  __syncthreads();
  for (int tok = warp; tok < FRT * FM; tok += 8) {
    const int r = tok / FM, m = tok % FM;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = sm.xh[r][m][lane + 32 * q];
    float mean, var;
    warp_ln_stats(v, mean, var);
    const float rstd = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = lane + 32 * q;
      v[q] = (v[q] - mean) * rstd * __ldg(P.pre_w + c) + __ldg(P.pre_b + c);       // Z (src/fusion.py:491)
    }
    warp_ln_stats(v, mean, var);                       // statistics shared by q_in_ln / k_in_ln / v_in_ln
    const float std2 = sqrtf(var + 1e-5f), rstd2 = 1.f / std2;
#pragma unroll
    for (int q = 0; q < 8; ++q) sm.xh[r][m][lane + 32 * q] = (v[q] - mean) * rstd2;
    if (lane == 0) { sm.zmean[r][m] = mean; sm.zstd[r][m] = std2; }
  }
  __syncthreads();

  // ---- phase 2: the two cross-modal attention layers ------------------------------------------------
  const int h = n >> 6;   // head of column n
#pragma unroll 1
  for (int i = 0; i < FM; ++i) {
    float q[FRT], kk[FRT][FM], vv[FRT][FM];
    {
      const float bq = __ldg(P.qkv_b[i][0] + n), bk = __ldg(P.qkv_b[i][1] + n), bv = __ldg(P.qkv_b[i][2] + n);
#pragma unroll
      for (int r = 0; r < FRT; ++r) {
        q[r] = bq;
#pragma unroll
        for (int m = 0; m < FM; ++m) { kk[r][m] = bk; vv[r][m] = bv; }
      }
    }
    const float *wq = P.qkv_wt[i][0] + n, *wk = P.qkv_wt[i][1] + n, *wv = P.qkv_wt[i][2] + n;
#pragma unroll 1
    for (int k = 0; k < FD; k += 4) {
      float cq[4], ck[4], cv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cq[u] = __ldg(wq + (size_t)(k + u) * FD);
        ck[u] = __ldg(wk + (size_t)(k + u) * FD);
        cv[u] = __ldg(wv + (size_t)(k + u) * FD);
      }
#pragma unroll
      for (int r = 0; r < FRT; ++r) {
#pragma unroll
        for (int m = 0; m < FM; ++m) {
          const float4 x = *reinterpret_cast<const float4*>(&sm.xh[r][m][k]);
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            kk[r][m] = fmaf(ck[u], xs[u], kk[r][m]);
            vv[r][m] = fmaf(cv[u], xs[u], vv[r][m]);
            if (m == i) q[r] = fmaf(cq[u], xs[u], q[r]);
          }
        }
      }
    }
    // scores[r][h][key] = sum over the 64 columns of head h of q*k, / sqrt(64)   (src/attention.py:112)
#pragma unroll
    for (int r = 0; r < FRT; ++r)
#pragma unroll
      for (int m = 0; m < FM; ++m) {
        const float s = warp_sum(q[r] * kk[r][m]);
        if (lane == 0) sm.part[warp][r][m] = s;
      }
    __syncthreads();
    if (n < FRT * FHEADS) {
      const int r = n >> 2, hh = n & 3;
      long long row = row0 + r;
      const bool live = row < n_rows;
      if (!live) row = n_rows - 1;
      const int b = (int)(row / P.T), t = (int)(row % P.T);
      float s[FM];
      bool inv[FM], all_inv = true;
#pragma unroll
      for (int m = 0; m < FM; ++m) {
        inv[m] = P.mask ? (__ldg(P.mask + (size_t)b * FM + m) <= 0.f) : false;       // src/fusion.py:474
        all_inv = all_inv && inv[m];
        s[m] = (sm.part[2 * hh][r][m] + sm.part[2 * hh + 1][r][m]) * 0.125f;
        if (inv[m]) s[m] = -1e4f;                                                    // src/attention.py:115-118
      }
      const float mx = fmaxf(s[0], s[1]);
      float e0 = expf(s[0] - mx), e1 = expf(s[1] - mx);
      const float den = e0 + e1;
      float a[FM] = {e0 / den, e1 / den};
      if (all_inv) { a[0] = 0.f; a[1] = 0.f; }                                       // :124-128
      if (P.ds.thr != 0u) {                                                          // :130, site ATTN + i
        const Philox4 w = philox4x32_10((uint32_t)(hh >> 1), ds_clip(P.ds, b),
                                        P.ds.t0 + (uint32_t)t, AVHOT_SITE_ATTN + i, P.ds.k0, P.ds.k1);
#pragma unroll
        for (int m = 0; m < FM; ++m) {
          const uint32_t word = philox_word(w, (uint32_t)((hh & 1) * 2 + m));
          a[m] = word >= P.ds.thr ? a[m] * P.ds.scale : 0.f;
        }
      }
#pragma unroll
      for (int m = 0; m < FM; ++m) {
        sm.attn[r][hh][m] = a[m];
        if (P.attn && live) P.attn[(((size_t)i * n_rows + row) * FHEADS + hh) * FM + m] = a[m];
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < FRT; ++r)
      sm.buf[r][n] = sm.attn[r][h][0] * vv[r][0] + sm.attn[r][h][1] * vv[r][1];      // context, :133-136
    __syncthreads();
    // out_proj + residual (Z_i = x^ * std + mean) -> raw into hatt[:, i]
    {
      float acc[FRT];
      const float bo = __ldg(P.o_b[i] + n);
#pragma unroll
      for (int r = 0; r < FRT; ++r) acc[r] = bo;
      const float* wo = P.o_wt[i] + n;
#pragma unroll 2
      for (int k = 0; k < FD; k += 4) {
        float co[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) co[u] = __ldg(wo + (size_t)(k + u) * FD);
#pragma unroll
        for (int r = 0; r < FRT; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(&sm.buf[r][k]);
          acc[r] = fmaf(co[0], x.x, acc[r]); acc[r] = fmaf(co[1], x.y, acc[r]);
          acc[r] = fmaf(co[2], x.z, acc[r]); acc[r] = fmaf(co[3], x.w, acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < FRT; ++r)
        sm.hatt[r][i][n] = acc[r] + fmaf(sm.xh[r][i][n], sm.zstd[r][i], sm.zmean[r][i]);
    }
    __syncthreads();
    for (int r = warp; r < FRT; r += 8) {                                            // post_ln, src/fusion.py:502
      float v[8];
#pragma unroll
      for (int qd = 0; qd < 8; ++qd) v[qd] = sm.hatt[r][i][lane + 32 * qd];
      float mean, var;
      warp_ln_stats(v, mean, var);
      const float rstd = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
      for (int qd = 0; qd < 8; ++qd) {
        const int c = lane + 32 * qd;
        const float y = (v[qd] - mean) * rstd * __ldg(P.post_w + c) + __ldg(P.post_b + c);
        sm.hatt[r][i][c] = y;
        if (P.h_att && row0 + r < n_rows) P.h_att[((size_t)(row0 + r) * FM + i) * FD + c] = y;
      }
    }
    __syncthreads();   // also protects part / attn / buf for the next layer
  }

  // ---- phase 3: fusion gate, modality softmax, classifier ------------------------------------------------
  {
    const int g = n & (FG - 1), m = n >> 7;
    float acc[FRT];
    const float gb = __ldg(P.g0_b + g);
#pragma unroll
    for (int r = 0; r < FRT; ++r) acc[r] = gb;
    const float* w0 = P.g0_wt + g;
#pragma unroll 2
    for (int k = 0; k < FD; k += 4) {
      float c0[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c0[u] = __ldg(w0 + (size_t)(k + u) * FG);
#pragma unroll
      for (int r = 0; r < FRT; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(&sm.hatt[r][m][k]);
        acc[r] = fmaf(c0[0], x.x, acc[r]); acc[r] = fmaf(c0[1], x.y, acc[r]);
        acc[r] = fmaf(c0[2], x.z, acc[r]); acc[r] = fmaf(c0[3], x.w, acc[r]);
      }
    }
    const float w3 = __ldg(P.g3_w + g);
#pragma unroll
    for (int r = 0; r < FRT; ++r) {
      float v = fmaxf(acc[r], 0.f);                                                  // ReLU, src/fusion.py:447
      if (P.ds.thr != 0u) {                                                          // Dropout :448, site GATE
        long long row = row0 + r;
        if (row >= n_rows) row = n_rows - 1;
        const int b = (int)(row / P.T), t = (int)(row % P.T);
        const uint32_t e = (uint32_t)(m * FG + g);
        const Philox4 w = philox4x32_10(e >> 2, ds_clip(P.ds, b), P.ds.t0 + (uint32_t)t,
                                        AVHOT_SITE_GATE, P.ds.k0, P.ds.k1);
        v = philox_word(w, e & 3) >= P.ds.thr ? v * P.ds.scale : 0.f;
      }
      const float s = warp_sum(v * w3);
      if (lane == 0) sm.part[warp][r][0] = s;
    }
  }
  __syncthreads();
  if (n < FRT) {
    const int r = n;
    long long row = row0 + r;
    const bool live = row < n_rows;
    if (!live) row = n_rows - 1;
    const int b = (int)(row / P.T);
    const float gb3 = __ldg(P.g3_b);
    float s[FM];
#pragma unroll
    for (int m = 0; m < FM; ++m) {
      s[m] = sm.part[4 * m][r][0] + sm.part[4 * m + 1][r][0] + sm.part[4 * m + 2][r][0] + sm.part[4 * m + 3][r][0] + gb3;
      const bool inv = P.mask ? (__ldg(P.mask + (size_t)b * FM + m) <= 0.f) : false;
      if (inv) s[m] = -1e4f;                                                         // :513
    }
    const float mx = fmaxf(s[0], s[1]);
    const float e0 = expf(s[0] - mx), e1 = expf(s[1] - mx);
    float w0 = e0 / (e0 + e1), w1 = e1 / (e0 + e1);                                  // :514
    if (!isfinite(w0)) w0 = 0.f;                                                     // :516
    if (!isfinite(w1)) w1 = 0.f;
    const float den = fmaxf(w0 + w1, 1e-8f);                                         // :517
    w0 /= den; w1 /= den;                                                            // :518
    sm.fw[r][0] = w0; sm.fw[r][1] = w1;
    if (P.fusion_weights && live) {
      P.fusion_weights[(size_t)row * FM] = w0;
      P.fusion_weights[(size_t)row * FM + 1] = w1;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < FRT; ++r)
    sm.buf[r][n] = sm.fw[r][0] * sm.hatt[r][0][n] + sm.fw[r][1] * sm.hatt[r][1][n];  // :521
  __syncthreads();
  {
    const int r = n >> 4, c = n & 15;
    if (c < P.C && row0 + r < n_rows) {
      const float4* w = reinterpret_cast<const float4*>(P.cls_w + (size_t)c * FD);
      const float4* x = reinterpret_cast<const float4*>(&sm.buf[r][0]);
      float acc = 0.f;
#pragma unroll 8
      for (int k = 0; k < FD / 4; ++k) {
        const float4 a = __ldg(w + k), bq = x[k];
        acc = fmaf(a.x, bq.x, acc); acc = fmaf(a.y, bq.y, acc);
        acc = fmaf(a.z, bq.z, acc); acc = fmaf(a.w, bq.w, acc);
      }
      P.logits[(size_t)(row0 + r) * P.C + c] = acc + __ldg(P.cls_b + c);             // :522
    }
  }
}

// =====================================================================================================================
// Hoisted variant (token 0 pass-invariant: z[0] has one row per clip).  Under MC-dropout the audio
// embedding is the same in every pass, and so is everything computed from the audio token alone before the first
// stochastic op: Z_a = pre_ln(proj_a(z_a)), and its projections Q_0 (query of layer 0), K_0a, V_0a, K_1a, V_1a.
// fusion_clip_kernel computes those once per clip into clipbuf (B, 6, 256); fusion_pass_kernel then does only the
// per-pass work for rows (clip, pass): 5 instead of 10 D x D projections of x^, same results as hybrid_fusion_kernel.
// =====================================================================================================================
constexpr int FCB = 6;   // clipbuf slots: Z_a, Q0, K0a, V0a, K1a, V1a

__global__ void __launch_bounds__(FTHREADS, 2) fusion_clip_kernel(const FusionParams P, float* __restrict__ clipbuf) {
  __shared__ __align__(16) float xh[FRT][FD];
  const int n = threadIdx.x, lane = n & 31, warp = n >> 5;
  const int b0 = blockIdx.x * FRT;
  {
    float acc[FRT];
#pragma unroll
    for (int r = 0; r < FRT; ++r) acc[r] = 0.f;
    const float* src[FRT];
#pragma unroll
    for (int r = 0; r < FRT; ++r) src[r] = P.z[0] + (size_t)min(b0 + r, P.B - 1) * P.dim[0];
    const float* wt = P.proj_wt[0] + n;
    for (int k = 0; k < P.dim[0]; k += 4) {
      const float w0 = __ldg(wt + (size_t)k * FD), w1 = __ldg(wt + (size_t)(k + 1) * FD);
      const float w2 = __ldg(wt + (size_t)(k + 2) * FD), w3 = __ldg(wt + (size_t)(k + 3) * FD);
#pragma unroll
      for (int r = 0; r < FRT; ++r) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(src[r] + k));
        acc[r] = fmaf(w0, x.x, acc[r]); acc[r] = fmaf(w1, x.y, acc[r]);
        acc[r] = fmaf(w2, x.z, acc[r]); acc[r] = fmaf(w3, x.w, acc[r]);
      }
    }
    const float pb = __ldg(P.proj_b[0] + n);
#pragma unroll
    for (int r = 0; r < FRT; ++r) xh[r][n] = acc[r] + pb;
  }
  __syncthreads();
  for (int r = warp; r < FRT; r += 8) {
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = xh[r][lane + 32 * q];
    float mean, var;
    warp_ln_stats(v, mean, var);
    const float rstd = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = lane + 32 * q;
      v[q] = (v[q] - mean) * rstd * __ldg(P.pre_w + c) + __ldg(P.pre_b + c);
    }
    warp_ln_stats(v, mean, var);
    const float std2 = sqrtf(var + 1e-5f), rstd2 = 1.f / std2;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = lane + 32 * q;
      const float xn = (v[q] - mean) * rstd2;
      xh[r][c] = xn;
      // Z_a exactly as hybrid_fusion_kernel rebuilds it for the residual: x^ * std + mean
      if (b0 + r < P.B) clipbuf[((size_t)(b0 + r) * FCB + 0) * FD + c] = fmaf(xn, std2, mean);
    }
  }
  __syncthreads();
  // five projections of x^_a: Q0, K0a, V0a (layer 0 weights), K1a, V1a (layer 1 weights)
  const float* wptr[5] = {P.qkv_wt[0][0] + n, P.qkv_wt[0][1] + n, P.qkv_wt[0][2] + n, P.qkv_wt[1][1] + n, P.qkv_wt[1][2] + n};
  const float bsv[5] = {__ldg(P.qkv_b[0][0] + n), __ldg(P.qkv_b[0][1] + n), __ldg(P.qkv_b[0][2] + n),
                        __ldg(P.qkv_b[1][1] + n), __ldg(P.qkv_b[1][2] + n)};
  float acc[5][FRT];
#pragma unroll
  for (int j = 0; j < 5; ++j)
#pragma unroll
    for (int r = 0; r < FRT; ++r) acc[j][r] = bsv[j];
#pragma unroll 1
  for (int k = 0; k < FD; k += 4) {
    float c[5][4];
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
      for (int u = 0; u < 4; ++u) c[j][u] = __ldg(wptr[j] + (size_t)(k + u) * FD);
#pragma unroll
    for (int r = 0; r < FRT; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(&xh[r][k]);
      const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[j][r] = fmaf(c[j][u], xs[u], acc[j][r]);
    }
  }
#pragma unroll
  for (int r = 0; r < FRT; ++r)
    if (b0 + r < P.B)
#pragma unroll
      for (int j = 0; j < 5; ++j) clipbuf[((size_t)(b0 + r) * FCB + 1 + j) * FD + n] = acc[j][r];
}

constexpr int HRT = 32, HTHREADS = 512, HR = 8;   // rows per CTA, threads, rows per thread

struct FusionPassSmem {
  float xh[HRT][FD];           // x^_v
  float hatt[HRT][FM][FD];     // H_att
  float buf[HRT][FD];          // context, then fused
  float part[16][HRT];         // per-warp partial reductions (gate scores)
  float zmean[HRT], zstd[HRT], fw[HRT][FM];
};

// Thread = 2 adjacent output columns x 8 rows: every 128-bit broadcast load of x^ feeds 8 FMAs per weight matrix and
// the weights arrive as 64-bit coalesced loads.  Warp w = (row group w >> 2, head w & 3): the 32 lanes of a warp hold
// exactly the 64 columns of one head, so a head's QK^T dot product is a single warp reduction and the whole attention
// (softmax over the two modality tokens, masks, Philox dropout, PV) stays in registers.
__global__ void __launch_bounds__(HTHREADS, 1)
fusion_pass_kernel(const FusionParams P, const float* __restrict__ clipbuf) {
  extern __shared__ __align__(16) unsigned char fusion_pass_smem_raw[];
  FusionPassSmem& sm = *reinterpret_cast<FusionPassSmem*>(fusion_pass_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cp = tid & 127, n0 = 2 * cp, rg = tid >> 7, rbase = rg * HR;
  const int head = warp & 3;
  const long long n_rows = (long long)P.B * P.T;
  const long long row0 = (long long)blockIdx.x * HRT;
  int bclip[HR], tpass[HR];
#pragma unroll
  for (int j = 0; j < HR; ++j) {
    long long row = row0 + rbase + j;
    if (row >= n_rows) row = n_rows - 1;
    bclip[j] = (int)(row / P.T);
    tpass[j] = (int)(row % P.T);
  }
  auto cb = [&](int j, int slot) {      // this thread's two columns of a per-clip vector
    return __ldg(reinterpret_cast<const float2*>(clipbuf + ((size_t)bclip[j] * FCB + slot) * FD + n0));
  };

  // ---- phase 1: Z_v = pre_ln(proj_v(z_v)), x^_v ---------------------------------------------------------------
  {
    float a0[HR], a1[HR];
    const float* src[HR];
#pragma unroll
    for (int j = 0; j < HR; ++j) {
      src[j] = P.z[1] + (size_t)((long long)bclip[j] * P.rpc[1] + (P.rpc[1] == 1 ? 0 : tpass[j])) * P.dim[1];
      a0[j] = 0.f; a1[j] = 0.f;
    }
    const float* wt = P.proj_wt[1] + n0;
    for (int k = 0; k < P.dim[1]; k += 4) {
      float2 w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = __ldg(reinterpret_cast<const float2*>(wt + (size_t)(k + u) * FD));
#pragma unroll
      for (int j = 0; j < HR; ++j) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(src[j] + k));
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) { a0[j] = fmaf(w[u].x, xs[u], a0[j]); a1[j] = fmaf(w[u].y, xs[u], a1[j]); }
      }
    }
    const float2 pb = __ldg(reinterpret_cast<const float2*>(P.proj_b[1] + n0));
#pragma unroll
    for (int j = 0; j < HR; ++j)
      *reinterpret_cast<float2*>(&sm.xh[rbase + j][n0]) = make_float2(a0[j] + pb.x, a1[j] + pb.y);
  }
  __syncthreads();
  for (int r = warp; r < HRT; r += 16) {
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = sm.xh[r][lane + 32 * q];
    float mean, var;
    warp_ln_stats(v, mean, var);
    const float rstd = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = lane + 32 * q;
      v[q] = (v[q] - mean) * rstd * __ldg(P.pre_w + c) + __ldg(P.pre_b + c);
    }
    warp_ln_stats(v, mean, var);
    const float std2 = sqrtf(var + 1e-5f), rstd2 = 1.f / std2;
#pragma unroll
    for (int q = 0; q < 8; ++q) sm.xh[r][lane + 32 * q] = (v[q] - mean) * rstd2;
    if (lane == 0) { sm.zmean[r] = mean; sm.zstd[r] = std2; }
  }
  __syncthreads();

  // ---- phase 2: the two attention layers; layer i's query is token i ----------------------------------------------
#pragma unroll 1
  for (int i = 0; i < FM; ++i) {
    float q0[HR], q1[HR], k0[HR], k1[HR], v0[HR], v1[HR];
    {
      const float2 bq = __ldg(reinterpret_cast<const float2*>(P.qkv_b[i][0] + n0));
      const float2 bk = __ldg(reinterpret_cast<const float2*>(P.qkv_b[i][1] + n0));
      const float2 bv = __ldg(reinterpret_cast<const float2*>(P.qkv_b[i][2] + n0));
#pragma unroll
      for (int j = 0; j < HR; ++j) { q0[j] = bq.x; q1[j] = bq.y; k0[j] = bk.x; k1[j] = bk.y; v0[j] = bv.x; v1[j] = bv.y; }
    }
    const float *wq = P.qkv_wt[i][0] + n0, *wk = P.qkv_wt[i][1] + n0, *wv = P.qkv_wt[i][2] + n0;
    if (i == 0) {
#pragma unroll 2
      for (int k = 0; k < FD; k += 4) {
        float2 ck[4], cv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ck[u] = __ldg(reinterpret_cast<const float2*>(wk + (size_t)(k + u) * FD));
          cv[u] = __ldg(reinterpret_cast<const float2*>(wv + (size_t)(k + u) * FD));
        }
#pragma unroll
        for (int j = 0; j < HR; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(&sm.xh[rbase + j][k]);
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            k0[j] = fmaf(ck[u].x, xs[u], k0[j]); k1[j] = fmaf(ck[u].y, xs[u], k1[j]);
            v0[j] = fmaf(cv[u].x, xs[u], v0[j]); v1[j] = fmaf(cv[u].y, xs[u], v1[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < HR; ++j) { const float2 t2 = cb(j, 1); q0[j] = t2.x; q1[j] = t2.y; }   // Q_0: per clip
    } else {
#pragma unroll 1
      for (int k = 0; k < FD; k += 4) {
        float2 cq[4], ck[4], cv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          cq[u] = __ldg(reinterpret_cast<const float2*>(wq + (size_t)(k + u) * FD));
          ck[u] = __ldg(reinterpret_cast<const float2*>(wk + (size_t)(k + u) * FD));
          cv[u] = __ldg(reinterpret_cast<const float2*>(wv + (size_t)(k + u) * FD));
        }
#pragma unroll
        for (int j = 0; j < HR; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(&sm.xh[rbase + j][k]);
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            q0[j] = fmaf(cq[u].x, xs[u], q0[j]); q1[j] = fmaf(cq[u].y, xs[u], q1[j]);
            k0[j] = fmaf(ck[u].x, xs[u], k0[j]); k1[j] = fmaf(ck[u].y, xs[u], k1[j]);
            v0[j] = fmaf(cv[u].x, xs[u], v0[j]); v1[j] = fmaf(cv[u].y, xs[u], v1[j]);
          }
        }
      }
    }
    // attention of this warp's head for its 8 rows, entirely in registers
    const int ka = i == 0 ? 2 : 4, va = i == 0 ? 3 : 5;        // clipbuf slots of K_ia, V_ia
#pragma unroll
    for (int j = 0; j < HR; ++j) {
      const float2 kaud = cb(j, ka);
      float s0 = warp_sum(q0[j] * kaud.x + q1[j] * kaud.y) * 0.125f;      // key = audio token
      float s1 = warp_sum(q0[j] * k0[j] + q1[j] * k1[j]) * 0.125f;        // key = video token
      const int b = bclip[j];
      const bool inv0 = P.mask ? (__ldg(P.mask + (size_t)b * FM) <= 0.f) : false;
      const bool inv1 = P.mask ? (__ldg(P.mask + (size_t)b * FM + 1) <= 0.f) : false;
      if (inv0) s0 = -1e4f;
      if (inv1) s1 = -1e4f;
      const float mx = fmaxf(s0, s1);
      const float e0 = expf(s0 - mx), e1 = expf(s1 - mx);
      const float den = e0 + e1;
      float a0 = e0 / den, a1 = e1 / den;
      if (inv0 && inv1) { a0 = 0.f; a1 = 0.f; }
      if (P.ds.thr != 0u) {
        const Philox4 w = philox4x32_10((uint32_t)(head >> 1), ds_clip(P.ds, b),
                                        P.ds.t0 + (uint32_t)tpass[j], AVHOT_SITE_ATTN + i, P.ds.k0, P.ds.k1);
        a0 = philox_word(w, (uint32_t)((head & 1) * 2)) >= P.ds.thr ? a0 * P.ds.scale : 0.f;
        a1 = philox_word(w, (uint32_t)((head & 1) * 2 + 1)) >= P.ds.thr ? a1 * P.ds.scale : 0.f;
      }
      const float2 vaud = cb(j, va);
      *reinterpret_cast<float2*>(&sm.buf[rbase + j][n0]) = make_float2(a0 * vaud.x + a1 * v0[j], a0 * vaud.y + a1 * v1[j]);
    }
    __syncthreads();
    {
      float a0[HR], a1[HR];
      const float2 bo = __ldg(reinterpret_cast<const float2*>(P.o_b[i] + n0));
#pragma unroll
      for (int j = 0; j < HR; ++j) { a0[j] = bo.x; a1[j] = bo.y; }
      const float* wo = P.o_wt[i] + n0;
#pragma unroll 2
      for (int k = 0; k < FD; k += 4) {
        float2 co[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) co[u] = __ldg(reinterpret_cast<const float2*>(wo + (size_t)(k + u) * FD));
#pragma unroll
        for (int j = 0; j < HR; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(&sm.buf[rbase + j][k]);
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) { a0[j] = fmaf(co[u].x, xs[u], a0[j]); a1[j] = fmaf(co[u].y, xs[u], a1[j]); }
        }
      }
#pragma unroll
      for (int j = 0; j < HR; ++j) {
        const int r = rbase + j;
        float2 res;
        if (i == 0) {
          res = cb(j, 0);                                                         // Z_a, per clip
        } else {
          const float2 xv = *reinterpret_cast<const float2*>(&sm.xh[r][n0]);
          res = make_float2(fmaf(xv.x, sm.zstd[r], sm.zmean[r]), fmaf(xv.y, sm.zstd[r], sm.zmean[r]));
        }
        *reinterpret_cast<float2*>(&sm.hatt[r][i][n0]) = make_float2(a0[j] + res.x, a1[j] + res.y);
      }
    }
    __syncthreads();
    for (int r = warp; r < HRT; r += 16) {
      float v[8];
#pragma unroll
      for (int qd = 0; qd < 8; ++qd) v[qd] = sm.hatt[r][i][lane + 32 * qd];
      float mean, var;
      warp_ln_stats(v, mean, var);
      const float rstd = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
      for (int qd = 0; qd < 8; ++qd) {
        const int c = lane + 32 * qd;
        sm.hatt[r][i][c] = (v[qd] - mean) * rstd * __ldg(P.post_w + c) + __ldg(P.post_b + c);
      }
    }
    __syncthreads();
  }

  // ---- phase 3: gate, modality softmax, classifier ------------------------------------------------------------------
  {
    const int g0 = 2 * (tid & 63), m = (tid >> 6) & 1;      // warps 4rg + 2m, 4rg + 2m + 1 cover the 128 gate units
    float a0[HR], a1[HR];
    const float2 gb = __ldg(reinterpret_cast<const float2*>(P.g0_b + g0));
#pragma unroll
    for (int j = 0; j < HR; ++j) { a0[j] = gb.x; a1[j] = gb.y; }
    const float* w0 = P.g0_wt + g0;
#pragma unroll 2
    for (int k = 0; k < FD; k += 4) {
      float2 c0[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c0[u] = __ldg(reinterpret_cast<const float2*>(w0 + (size_t)(k + u) * FG));
#pragma unroll
      for (int j = 0; j < HR; ++j) {
        const float4 x = *reinterpret_cast<const float4*>(&sm.hatt[rbase + j][m][k]);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) { a0[j] = fmaf(c0[u].x, xs[u], a0[j]); a1[j] = fmaf(c0[u].y, xs[u], a1[j]); }
      }
    }
    const float2 w3 = __ldg(reinterpret_cast<const float2*>(P.g3_w + g0));
#pragma unroll
    for (int j = 0; j < HR; ++j) {
      float va = fmaxf(a0[j], 0.f), vb = fmaxf(a1[j], 0.f);
      if (P.ds.thr != 0u) {
        const uint32_t e = (uint32_t)(m * FG + g0);           // even: both elements are in one Philox block
        const Philox4 w = philox4x32_10(e >> 2, ds_clip(P.ds, bclip[j]),
                                        P.ds.t0 + (uint32_t)tpass[j], AVHOT_SITE_GATE, P.ds.k0, P.ds.k1);
        va = philox_word(w, e & 3) >= P.ds.thr ? va * P.ds.scale : 0.f;
        vb = philox_word(w, (e & 3) + 1) >= P.ds.thr ? vb * P.ds.scale : 0.f;
      }
      const float sres = warp_sum(va * w3.x + vb * w3.y);
      if (lane == 0) sm.part[warp][rbase + j] = sres;
    }
  }
  __syncthreads();
  if (tid < HRT) {
    const int r = tid;
    long long row = row0 + r;
    if (row >= n_rows) row = n_rows - 1;
    const int b = (int)(row / P.T);
    const float gb3 = __ldg(P.g3_b);
    const int wbase = 4 * (r / HR);            // warps of this row's group: modality m uses warps wbase + 2m, + 2m + 1
    float s[FM];
#pragma unroll
    for (int m = 0; m < FM; ++m) {
      s[m] = sm.part[wbase + 2 * m][r] + sm.part[wbase + 2 * m + 1][r] + gb3;
      const bool inv = P.mask ? (__ldg(P.mask + (size_t)b * FM + m) <= 0.f) : false;
      if (inv) s[m] = -1e4f;
    }
    const float mx = fmaxf(s[0], s[1]);
    const float e0 = expf(s[0] - mx), e1 = expf(s[1] - mx);
    float w0 = e0 / (e0 + e1), w1 = e1 / (e0 + e1);
    if (!isfinite(w0)) w0 = 0.f;
    if (!isfinite(w1)) w1 = 0.f;
    const float den = fmaxf(w0 + w1, 1e-8f);
    sm.fw[r][0] = w0 / den;
    sm.fw[r][1] = w1 / den;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < HR; ++j) {
    const int r = rbase + j;
    const float2 h0 = *reinterpret_cast<const float2*>(&sm.hatt[r][0][n0]);
    const float2 h1 = *reinterpret_cast<const float2*>(&sm.hatt[r][1][n0]);
    *reinterpret_cast<float2*>(&sm.buf[r][n0]) =
        make_float2(sm.fw[r][0] * h0.x + sm.fw[r][1] * h1.x, sm.fw[r][0] * h0.y + sm.fw[r][1] * h1.y);
  }
  __syncthreads();
  {
    const int r = tid >> 4, c = tid & 15;
    if (c < P.C && row0 + r < n_rows) {
      const float4* w = reinterpret_cast<const float4*>(P.cls_w + (size_t)c * FD);
      const float4* x = reinterpret_cast<const float4*>(&sm.buf[r][0]);
      float acc = 0.f;
#pragma unroll 8
      for (int k = 0; k < FD / 4; ++k) {
        const float4 a = __ldg(w + k), bq = x[k];
        acc = fmaf(a.x, bq.x, acc); acc = fmaf(a.y, bq.y, acc);
        acc = fmaf(a.z, bq.z, acc); acc = fmaf(a.w, bq.w, acc);
      }
      P.logits[(size_t)(row0 + r) * P.C + c] = acc + __ldg(P.cls_b + c);
    }
  }
}

// ---- live baseline head: cat -> Linear -> ReLU -> Linear  (src/train.py:81-85, :153-164) -----------------
struct HeadParams {
  const float* z[FM];
  int rpc[FM], dim[FM];
  float* logits;
  int B, T, C, in_dim;
  const float *w0t, *b0, *w2, *b2;
};

__global__ void __launch_bounds__(FTHREADS) concat_head_kernel(const HeadParams P) {
  extern __shared__ __align__(16) unsigned char head_smem_raw[];
  float* xin = reinterpret_cast<float*>(head_smem_raw);          // [FRT][in_dim]
  float* hid = xin + (size_t)FRT * P.in_dim;                     // [FRT][FD]
  const int n = threadIdx.x;
  const long long n_rows = (long long)P.B * P.T;
  const long long row0 = (long long)blockIdx.x * FRT;
  for (int idx = n; idx < FRT * P.in_dim; idx += FTHREADS) {
    const int r = idx / P.in_dim;
    int k = idx % P.in_dim;
    long long row = row0 + r;
    if (row >= n_rows) row = n_rows - 1;
    const long long b = row / P.T, t = row % P.T;
    int m = 0;
    while (m + 1 < FM && k >= P.dim[m]) { k -= P.dim[m]; ++m; }
    xin[idx] = __ldg(P.z[m] + (size_t)(b * P.rpc[m] + (P.rpc[m] == 1 ? 0 : t)) * P.dim[m] + k);
  }
  __syncthreads();
  float acc[FRT];
  const float b0 = __ldg(P.b0 + n);
#pragma unroll
  for (int r = 0; r < FRT; ++r) acc[r] = b0;
  const float* w = P.w0t + n;
  for (int k = 0; k < P.in_dim; k += 4) {
    float c0[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) c0[u] = __ldg(w + (size_t)(k + u) * FD);
#pragma unroll
    for (int r = 0; r < FRT; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(xin + (size_t)r * P.in_dim + k);
      acc[r] = fmaf(c0[0], x.x, acc[r]); acc[r] = fmaf(c0[1], x.y, acc[r]);
      acc[r] = fmaf(c0[2], x.z, acc[r]); acc[r] = fmaf(c0[3], x.w, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < FRT; ++r) hid[r * FD + n] = fmaxf(acc[r], 0.f);
  __syncthreads();
  const int r = n >> 4, c = n & 15;
  if (c < P.C && row0 + r < n_rows) {
    const float4* wc = reinterpret_cast<const float4*>(P.w2 + (size_t)c * FD);
    const float4* x = reinterpret_cast<const float4*>(hid + r * FD);
    float a = 0.f;
#pragma unroll 8
    for (int k = 0; k < FD / 4; ++k) {
      const float4 u = __ldg(wc + k), v = x[k];
      a = fmaf(u.x, v.x, a); a = fmaf(u.y, v.y, a); a = fmaf(u.z, v.z, a); a = fmaf(u.w, v.w, a);
    }
    P.logits[(size_t)(row0 + r) * P.C + c] = a + __ldg(P.b2 + c);
  }
}

static size_t al(size_t n_floats) { return (n_floats * sizeof(float) + 255) & ~(size_t)255; }

}  // namespace avhot

using namespace avhot;

extern "C" int avhot_fusion_create(const avhot_fusion_desc* f, const avhot_head_desc* hd, avhot_fusion_t* out,
                                   avhot_stream_t stream) {
  AVHOT_REQUIRE(out && ((f != nullptr) != (hd != nullptr)), AVHOT_ERR_INVALID,
                "fusion_create: exactly one of fusion / head must be given");
  cudaStream_t s = as_stream(stream);
  avhot_fusion* h = new avhot_fusion();
  memset(h, 0, sizeof(*h));
  size_t total = 0;
  if (f) {
    if (!(f->num_modalities == FM && f->hidden_dim == FD && f->num_heads == FHEADS && f->gate_hidden == FG &&
          f->num_classes >= 1 && f->num_classes <= FCMAX)) {
      set_error("fusion_create: this build is specialised for 2 modalities, hidden_dim 256, 4 heads, gate 128, "
                "<= 16 classes (got M=%d D=%d heads=%d G=%d C=%d)", f->num_modalities, f->hidden_dim, f->num_heads,
                f->gate_hidden, f->num_classes);
      delete h;
      return AVHOT_ERR_UNSUPPORTED;
    }
    for (int m = 0; m < FM; ++m) {
      if (f->modality_dim[m] <= 0 || (f->modality_dim[m] & 3)) {
        set_error("fusion_create: modality_dim[%d]=%d must be a positive multiple of 4", m, f->modality_dim[m]);
        delete h;
        return AVHOT_ERR_UNSUPPORTED;
      }
      total += al((size_t)f->modality_dim[m] * FD) + al(FD);
      total += 4 * (al((size_t)FD * FD) + al(FD));
    }
    total += 4 * al(FD) + al((size_t)FD * FG) + al(FG) + al(FG) + al(1) + al((size_t)f->num_classes * FD) +
             al(f->num_classes);
    total += al(fusion_tc_packed_bytes() / sizeof(float)) + al(FUS_VEC_FLOATS);
  } else {
    if (!(hd->hidden_dim == FD && hd->num_classes >= 1 && hd->num_classes <= FCMAX && hd->in_dim > 0 &&
          (hd->in_dim & 3) == 0 && hd->in_dim <= 2048)) {
      set_error("fusion_create(head): hidden_dim must be 256, classes <= 16, in_dim a multiple of 4 <= 2048 "
                "(got %d, %d, %d)", hd->hidden_dim, hd->num_classes, hd->in_dim);
      delete h;
      return AVHOT_ERR_UNSUPPORTED;
    }
    total += al((size_t)hd->in_dim * FD) + al(FD) + al((size_t)hd->num_classes * FD) + al(hd->num_classes);
  }
  cudaError_t e = cudaMalloc(&h->arena, total);
  if (e != cudaSuccess) {
    set_error("fusion_create: cudaMalloc(%zu) -> %s", total, cudaGetErrorString(e));
    delete h;
    return AVHOT_ERR_CUDA;
  }
  char* p = reinterpret_cast<char*>(h->arena);
  auto take = [&](size_t n_floats) { float* r = reinterpret_cast<float*>(p); p += al(n_floats); return r; };
  bool ok = true;
  auto copy = [&](float* dst, const float* src, size_t n_floats) {
    if (!src) { ok = false; return; }
    if (cudaMemcpyAsync(dst, src, n_floats * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess) ok = false;
  };
  auto fold = [&](const float* W, const float* b, const float* lw, const float* lb, float* Wt, float* bf, int N,
                  int K) {
    if (!W || !b) { ok = false; return; }
    fold_ln_kernel<<<N, 128, 0, s>>>(W, b, lw, lb, Wt, bf, N, K);
    if (cudaGetLastError() != cudaSuccess) ok = false;
  };
  if (f) {
    h->kind = 0; h->M = FM; h->D = FD; h->heads = FHEADS; h->C = f->num_classes; h->G = FG;
    for (int m = 0; m < FM; ++m) {
      h->dim[m] = f->modality_dim[m];
      h->proj_wt[m] = take((size_t)h->dim[m] * FD);
      h->proj_b[m] = take(FD);
      fold(f->proj_w[m], f->proj_b[m], nullptr, nullptr, h->proj_wt[m], h->proj_b[m], FD, h->dim[m]);
      const avhot_xattn_desc& a = f->attn[m];
      if (a.query_dim != FD || a.key_dim != FD || a.hidden_dim != FD || a.num_heads != FHEADS) ok = false;
      const float* W[3] = {a.q_w, a.k_w, a.v_w};
      const float* bb[3] = {a.q_b, a.k_b, a.v_b};
      const float* lw[3] = {a.q_ln_w, a.k_ln_w, a.v_ln_w};
      const float* lb[3] = {a.q_ln_b, a.k_ln_b, a.v_ln_b};
      for (int j = 0; j < 3; ++j) {
        h->qkv_wt[m][j] = take((size_t)FD * FD);
        h->qkv_b[m][j] = take(FD);
        if (!lw[j] || !lb[j]) ok = false;
        fold(W[j], bb[j], lw[j], lb[j], h->qkv_wt[m][j], h->qkv_b[m][j], FD, FD);
      }
      h->o_wt[m] = take((size_t)FD * FD);
      h->o_b[m] = take(FD);
      fold(a.o_w, a.o_b, nullptr, nullptr, h->o_wt[m], h->o_b[m], FD, FD);
    }
    h->pre_w = take(FD); h->pre_b = take(FD); h->post_w = take(FD); h->post_b = take(FD);
    copy(h->pre_w, f->pre_ln_w, FD); copy(h->pre_b, f->pre_ln_b, FD);
    copy(h->post_w, f->post_ln_w, FD); copy(h->post_b, f->post_ln_b, FD);
    h->g0_wt = take((size_t)FD * FG); h->g0_b = take(FG);
    fold(f->gate0_w, f->gate0_b, nullptr, nullptr, h->g0_wt, h->g0_b, FG, FD);
    h->g3_w = take(FG); h->g3_b = take(1);
    copy(h->g3_w, f->gate3_w, FG); copy(h->g3_b, f->gate3_b, 1);
    h->cls_w = take((size_t)h->C * FD); h->cls_b = take(h->C);
    copy(h->cls_w, f->cls_w, (size_t)h->C * FD); copy(h->cls_b, f->cls_b, h->C);
    // tensor-core per-pass kernel: packed fp16 weights (from the folded k-major fp32 copies above) + small vectors
    h->tc_w = take(fusion_tc_packed_bytes() / sizeof(float));
    h->tc_vec = take(FUS_VEC_FLOATS);
    if (ok && h->dim[1] % 64 == 0 && h->dim[1] <= 256 && h->C <= 8) {
      auto pack = [&](const float* src_t, int row0, int N, int K) {
        if (fusion_tc_pack_rows(src_t, h->tc_w, row0, N, K, s) != AVHOT_OK) ok = false;
      };
      pack(h->proj_wt[1], 0, FD, h->dim[1]);
      pack(h->qkv_wt[0][1], 256, FD, FD);  pack(h->qkv_wt[0][2], 512, FD, FD);
      pack(h->qkv_wt[1][0], 768, FD, FD);  pack(h->qkv_wt[1][1], 1024, FD, FD);  pack(h->qkv_wt[1][2], 1280, FD, FD);
      pack(h->o_wt[0], 1536, FD, FD);      pack(h->o_wt[1], 1792, FD, FD);
      pack(h->g0_wt, 2048, FG, FD);
      if (cudaMemsetAsync(h->tc_vec, 0, sizeof(float) * FUS_VEC_FLOATS, s) != cudaSuccess) ok = false;
      copy(h->tc_vec + VEC_PROJ_B, h->proj_b[1], FD);
      copy(h->tc_vec + VEC_PRE_W, h->pre_w, FD);   copy(h->tc_vec + VEC_PRE_B, h->pre_b, FD);
      copy(h->tc_vec + VEC_POST_W, h->post_w, FD); copy(h->tc_vec + VEC_POST_B, h->post_b, FD);
      copy(h->tc_vec + VEC_BK0, h->qkv_b[0][1], FD); copy(h->tc_vec + VEC_BV0, h->qkv_b[0][2], FD);
      copy(h->tc_vec + VEC_BQ1, h->qkv_b[1][0], FD); copy(h->tc_vec + VEC_BK1, h->qkv_b[1][1], FD);
      copy(h->tc_vec + VEC_BV1, h->qkv_b[1][2], FD);
      copy(h->tc_vec + VEC_BO0, h->o_b[0], FD);    copy(h->tc_vec + VEC_BO1, h->o_b[1], FD);
      copy(h->tc_vec + VEC_G0_B, h->g0_b, FG);     copy(h->tc_vec + VEC_G3_W, h->g3_w, FG);
      copy(h->tc_vec + VEC_G3_B, h->g3_b, 1);
      copy(h->tc_vec + VEC_CLS_W, h->cls_w, (size_t)h->C * FD);
      copy(h->tc_vec + VEC_CLS_B, h->cls_b, h->C);
    } else {
      h->tc_w = nullptr;
    }
  } else {
    h->kind = 1; h->M = FM; h->D = FD; h->C = hd->num_classes; h->head_in = hd->in_dim; h->head_hidden = FD;
    h->h0_wt = take((size_t)hd->in_dim * FD); h->h0_b = take(FD);
    fold(hd->w0, hd->b0, nullptr, nullptr, h->h0_wt, h->h0_b, FD, hd->in_dim);
    h->cls_w = take((size_t)h->C * FD); h->cls_b = take(h->C);
    copy(h->cls_w, hd->w2, (size_t)h->C * FD); copy(h->cls_b, hd->b2, h->C);
  }
  if (!ok) {
    set_error("fusion_create: missing tensor, inconsistent attention dims, or device copy failure (%s)",
              cudaGetErrorString(cudaGetLastError()));
    cudaFree(h->arena);
    delete h;
    return AVHOT_ERR_INVALID;
  }
  *out = h;
  return AVHOT_OK;
}

extern "C" void avhot_fusion_destroy(avhot_fusion_t h) {
  if (!h) return;
  cudaFree(h->arena);
  delete h;
}

extern "C" size_t avhot_fusion_workspace_bytes(avhot_fusion_t h, int B, int T) {
  (void)T;
  if (!h || h->kind != 0 || B <= 0) return 0;
  return sizeof(float) * (size_t)B * FCB * FD;
}

static int fusion_forward_impl(avhot_fusion_t h, const float* const* z_host_array,
                               const int* z_rows_per_clip_host, const float* mask, float* logits,
                               float* fusion_weights, float* attn, float* h_att, int B, int T, float p,
                               uint64_t seed, uint64_t clip0, const uint32_t* clip_ids, uint32_t t0, int precision,
                               void* workspace, size_t workspace_bytes, avhot_stream_t stream);

extern "C" int avhot_fusion_forward(avhot_fusion_t h, const float* const* z_host_array,
                                    const int* z_rows_per_clip_host, const float* mask, float* logits,
                                    float* fusion_weights, float* attn, float* h_att, int B, int T, float p,
                                    uint64_t seed, uint64_t clip0, uint32_t t0, int precision, void* workspace,
                                    size_t workspace_bytes, avhot_stream_t stream) {
  return fusion_forward_impl(h, z_host_array, z_rows_per_clip_host, mask, logits, fusion_weights, attn, h_att, B, T, p,
                             seed, clip0, nullptr, t0, precision, workspace, workspace_bytes, stream);
}

extern "C" int avhot_fusion_forward_ids(avhot_fusion_t h, const float* const* z_host_array,
                                        const int* z_rows_per_clip_host, const float* mask, float* logits, int B, int T,
                                        float p, uint64_t seed, const uint32_t* clip_ids, uint32_t t0, int precision,
                                        void* workspace, size_t workspace_bytes, avhot_stream_t stream) {
  AVHOT_REQUIRE(clip_ids, AVHOT_ERR_INVALID, "fusion_forward_ids: clip_ids is NULL");
  return fusion_forward_impl(h, z_host_array, z_rows_per_clip_host, mask, logits, nullptr, nullptr, nullptr, B, T, p,
                             seed, 0, clip_ids, t0, precision, workspace, workspace_bytes, stream);
}

static int fusion_forward_impl(avhot_fusion_t h, const float* const* z_host_array,
                               const int* z_rows_per_clip_host, const float* mask, float* logits,
                               float* fusion_weights, float* attn, float* h_att, int B, int T, float p,
                               uint64_t seed, uint64_t clip0, const uint32_t* clip_ids, uint32_t t0, int precision,
                               void* workspace, size_t workspace_bytes, avhot_stream_t stream) {
  AVHOT_REQUIRE(h && z_host_array && z_rows_per_clip_host && logits && B >= 0 && T >= 1, AVHOT_ERR_INVALID,
                "fusion_forward: bad arguments");
  AVHOT_REQUIRE(p >= 0.f && p < 1.f, AVHOT_ERR_INVALID, "fusion_forward: dropout p=%g outside [0, 1)", (double)p);
  AVHOT_REQUIRE(precision == AVHOT_PREC_FP32 || precision == AVHOT_PREC_TC, AVHOT_ERR_UNSUPPORTED,
                "fusion_forward: precision %d not built", precision);
  if (B == 0) return AVHOT_OK;
  const long long rows = (long long)B * T;
  const long long grid = (rows + FRT - 1) / FRT;
  AVHOT_REQUIRE(grid <= 0x7fffffffLL, AVHOT_ERR_INVALID, "fusion_forward: B*T too large for one call");
  for (int m = 0; m < FM; ++m) {
    AVHOT_REQUIRE(z_host_array[m], AVHOT_ERR_INVALID, "fusion_forward: z[%d] is NULL", m);
    AVHOT_REQUIRE(z_rows_per_clip_host[m] == 1 || z_rows_per_clip_host[m] == T, AVHOT_ERR_INVALID,
                  "fusion_forward: z_rows_per_clip[%d]=%d must be 1 or T=%d", m, z_rows_per_clip_host[m], T);
    AVHOT_REQUIRE((reinterpret_cast<uintptr_t>(z_host_array[m]) & 15) == 0, AVHOT_ERR_INVALID,
                  "fusion_forward: z[%d] must be 16-byte aligned", m);
  }
  cudaStream_t s = as_stream(stream);
  if (h->kind == 0) {
    FusionParams P;
    for (int m = 0; m < FM; ++m) {
      P.z[m] = z_host_array[m]; P.rpc[m] = z_rows_per_clip_host[m]; P.dim[m] = h->dim[m];
      P.proj_wt[m] = h->proj_wt[m]; P.proj_b[m] = h->proj_b[m];
      for (int j = 0; j < 3; ++j) { P.qkv_wt[m][j] = h->qkv_wt[m][j]; P.qkv_b[m][j] = h->qkv_b[m][j]; }
      P.o_wt[m] = h->o_wt[m]; P.o_b[m] = h->o_b[m];
    }
    P.mask = mask; P.logits = logits; P.fusion_weights = fusion_weights; P.attn = attn; P.h_att = h_att;
    P.B = B; P.T = T; P.C = h->C;
    P.pre_w = h->pre_w; P.pre_b = h->pre_b; P.post_w = h->post_w; P.post_b = h->post_b;
    P.g0_wt = h->g0_wt; P.g0_b = h->g0_b; P.g3_w = h->g3_w; P.g3_b = h->g3_b;
    P.cls_w = h->cls_w; P.cls_b = h->cls_b;
    P.ds = make_drop(p, seed, clip0, t0);
    P.ds.clip_ids = clip_ids;
    // token 0 given once per clip and no per-row diagnostics requested: hoist its pass-invariant half
    if (P.rpc[0] == 1 && !fusion_weights && !attn && !h_att && workspace &&
        workspace_bytes >= avhot_fusion_workspace_bytes(h, B, T)) {
      float* clipbuf = reinterpret_cast<float*>(workspace);
      {
        ProfScope prof("fusion_clip", s);
        fusion_clip_kernel<<<ceil_div(B, FRT), FTHREADS, 0, s>>>(P, clipbuf);
        AVHOT_LAUNCH_CHECK();
      }
      // (the tensor-core kernel reads the per-clip vectors with 128-bit loads: an unaligned caller workspace takes the fp32 pass kernel)
      if (precision == AVHOT_PREC_TC && h->tc_w && (reinterpret_cast<uintptr_t>(clipbuf) & 15) == 0) {
        FusTcParams Q;
        Q.z_v = P.z[1]; Q.rpc_v = P.rpc[1]; Q.dim_v = P.dim[1];
        Q.clipbuf = clipbuf; Q.mask = mask; Q.logits = logits;
        Q.B = B; Q.T = T; Q.C = h->C; Q.vec = h->tc_vec; Q.ds = P.ds; Q.trace = nullptr;
        return fusion_tc_forward(h->tc_w, Q, s);
      }
      const long long grid2 = (rows + HRT - 1) / HRT;
      AVHOT_CUDA(cudaFuncSetAttribute(fusion_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(FusionPassSmem)));
      ProfScope prof("fusion_pass", s);
      fusion_pass_kernel<<<(unsigned)grid2, HTHREADS, sizeof(FusionPassSmem), s>>>(P, clipbuf);
      AVHOT_LAUNCH_CHECK();
      return AVHOT_OK;
    }
    AVHOT_CUDA(cudaFuncSetAttribute(hybrid_fusion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(FusionSmem)));
    ProfScope prof("hybrid_fusion", s);
    hybrid_fusion_kernel<<<(unsigned)grid, FTHREADS, sizeof(FusionSmem), s>>>(P);
    AVHOT_LAUNCH_CHECK();
  } else {
    AVHOT_REQUIRE(!fusion_weights && !attn && !h_att, AVHOT_ERR_INVALID,
                  "fusion_forward: the concat head has no attention outputs");
    HeadParams P;
    int sum = 0;
    for (int m = 0; m < FM; ++m) { P.z[m] = z_host_array[m]; P.rpc[m] = z_rows_per_clip_host[m]; }
    // the head was created with in_dim only: the per-modality split is in_dim / M (equal output_dim encoders,
    // src/train.py:58-66)
    for (int m = 0; m < FM; ++m) { P.dim[m] = h->head_in / FM; sum += P.dim[m]; }
    AVHOT_REQUIRE(sum == h->head_in && (P.dim[0] & 3) == 0, AVHOT_ERR_UNSUPPORTED,
                  "fusion_forward(head): in_dim=%d must split evenly over %d modalities", h->head_in, FM);
    P.logits = logits; P.B = B; P.T = T; P.C = h->C; P.in_dim = h->head_in;
    P.w0t = h->h0_wt; P.b0 = h->h0_b; P.w2 = h->cls_w; P.b2 = h->cls_b;
    const size_t smem = sizeof(float) * FRT * (size_t)(h->head_in + FD);
    AVHOT_CUDA(cudaFuncSetAttribute(concat_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof("concat_head", s);
    concat_head_kernel<<<(unsigned)grid, FTHREADS, smem, s>>>(P);
    AVHOT_LAUNCH_CHECK();
  }
  return AVHOT_OK;
}
