// SURVEY.md 8(f) "next" rows built to the same bar as the hot path:
//   EarlyFusion  (src/fusion.py:19-138)   concat(mask_m * z_m) -> Linear-LN-ReLU-Dropout x2 -> Linear
//   LateFusion   (src/fusion.py:141-248)  per-modality Linear-ReLU-Dropout-Linear, softmax(fusion_logits) * mask,
//                                         renormalised, weighted sum of the per-modality logits
//   calibration  (src/uncertainty.py:88-183, :512-544)  confidence histogram (count, sum conf, correct) + NLL
// fp32 CUDA-core kernels in the style of fusion.cu: CTA = 16 rows (clip, pass), thread = hidden column, everything
// between the layers in shared memory, Philox dropout at the reference's nn.Dropout sites.
#include "common.cuh"

struct avhot_head {
  int kind;              // 0 early, 1 late, 2 late with uncertainty
  // late with uncertainty (hidden_dim D > 0: Linear(d, D)-ReLU-Linear(D, C | 1); D == 0: Linear(d, C | 1))
  float *uw0t[2][AVHOT_MAX_MODALITIES], *ub0[2][AVHOT_MAX_MODALITIES];   // [0] classifier, [1] uncertainty head
  float *uw3[2][AVHOT_MAX_MODALITIES], *ub3[2][AVHOT_MAX_MODALITIES];
  float eps;
  int M, D, C, in_dim;
  int dim[AVHOT_MAX_MODALITIES];
  // early
  float *w0t, *b0, *ln1w, *ln1b, *w4t, *b4, *ln5w, *ln5b, *w8, *b8;
  // late
  float *lw0t[AVHOT_MAX_MODALITIES], *lb0[AVHOT_MAX_MODALITIES], *lw3[AVHOT_MAX_MODALITIES], *lb3[AVHOT_MAX_MODALITIES];
  float* fusion_logits;
  float* arena;
};

namespace avhot {

constexpr int HD = 256, HRT16 = 16, HTH = 256, HCMAX = 16;

// Wt[k][n] = W[n][k]
__global__ void transpose_rows_kernel(const float* __restrict__ W, float* __restrict__ Wt, int N, int K) {
  const int n = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) Wt[(size_t)k * N + n] = W[(size_t)n * K + k];
}

struct HeadRow { long long row; int b, t; };
__device__ __forceinline__ HeadRow head_row(long long row0, int r, long long n_rows, int T) {
  long long row = row0 + r;
  if (row >= n_rows) row = n_rows - 1;
  return HeadRow{row, (int)(row / T), (int)(row % T)};
}

// LayerNorm + ReLU + dropout of the 16 rows in `h` (in place); lane owns columns 4l..4l+3 and 128+4l..128+4l+3,
// i.e. exactly two Philox blocks of the (hidden,) mask of its row
__device__ __forceinline__ void ln_relu_drop_rows(float (*h)[HD], const float* __restrict__ lw, const float* __restrict__ lb,
                                                  long long row0, long long n_rows, int T, uint32_t site, const DropSpec& ds) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = 4 * lane, c1 = 128 + 4 * lane;
  for (int r = warp; r < HRT16; r += HTH / 32) {
    const float4 a = *reinterpret_cast<const float4*>(&h[r][c0]), b = *reinterpret_cast<const float4*>(&h[r][c1]);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += v[q];
    const float mean = warp_sum(s) * (1.f / HD);
    float sq = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) { const float d = v[q] - mean; sq = fmaf(d, d, sq); }
    const float rstd = 1.f / sqrtf(warp_sum(sq) * (1.f / HD) + 1e-5f);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(lw + c0)), w1 = __ldg(reinterpret_cast<const float4*>(lw + c1));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(lb + c0)), b1 = __ldg(reinterpret_cast<const float4*>(lb + c1));
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = fmaxf((v[q] - mean) * rstd * wv[q] + bv[q], 0.f);
    if (ds.thr != 0u) {
      const HeadRow hr = head_row(row0, r, n_rows, T);
      const uint32_t g = (uint32_t)(ds.clip0 + (uint64_t)hr.b), pass = ds.t0 + (uint32_t)hr.t;
      const Philox4 p0 = philox4x32_10((uint32_t)lane, g, pass, site, ds.k0, ds.k1);
      const Philox4 p1 = philox4x32_10((uint32_t)(32 + lane), g, pass, site, ds.k0, ds.k1);
      const uint32_t wd[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = wd[q] >= ds.thr ? v[q] * ds.scale : 0.f;
    }
    *reinterpret_cast<float4*>(&h[r][c0]) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(&h[r][c1]) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

struct EarlyParams {
  const float* z[AVHOT_MAX_MODALITIES];
  int rpc[AVHOT_MAX_MODALITIES], dim[AVHOT_MAX_MODALITIES];
  const float* mask;
  float* logits;
  int M, B, T, C, in_dim;
  const float *w0t, *b0, *ln1w, *ln1b, *w4t, *b4, *ln5w, *ln5b, *w8, *b8;
  DropSpec ds;
};

__global__ void __launch_bounds__(HTH) early_fusion_kernel(const EarlyParams P) {
  extern __shared__ __align__(16) unsigned char early_smem_raw[];
  float* xin = reinterpret_cast<float*>(early_smem_raw);                         // [16][in_dim]
  float (*hA)[HD] = reinterpret_cast<float (*)[HD]>(xin + (size_t)HRT16 * P.in_dim);
  float (*hB)[HD] = hA + HRT16;
  const int n = threadIdx.x;
  const long long n_rows = (long long)P.B * P.T, row0 = (long long)blockIdx.x * HRT16;
  // concat of the masked modality features (src/fusion.py:113-137)
  for (int idx = n; idx < HRT16 * P.in_dim; idx += HTH) {
    const int r = idx / P.in_dim;
    int k = idx % P.in_dim, m = 0;
    while (m + 1 < P.M && k >= P.dim[m]) { k -= P.dim[m]; ++m; }
    const HeadRow hr = head_row(row0, r, n_rows, P.T);
    float x = __ldg(P.z[m] + (size_t)((long long)hr.b * P.rpc[m] + (P.rpc[m] == 1 ? 0 : hr.t)) * P.dim[m] + k);
    if (P.mask) x *= __ldg(P.mask + (size_t)hr.b * P.M + m);
    xin[idx] = x;
  }
  __syncthreads();
  {
    float acc[HRT16];
    const float b = __ldg(P.b0 + n);
#pragma unroll
    for (int r = 0; r < HRT16; ++r) acc[r] = b;
    const float* w = P.w0t + n;
    for (int k = 0; k < P.in_dim; k += 4) {
      float c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c[u] = __ldg(w + (size_t)(k + u) * HD);
#pragma unroll
      for (int r = 0; r < HRT16; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(xin + (size_t)r * P.in_dim + k);
        acc[r] = fmaf(c[0], x.x, acc[r]); acc[r] = fmaf(c[1], x.y, acc[r]);
        acc[r] = fmaf(c[2], x.z, acc[r]); acc[r] = fmaf(c[3], x.w, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < HRT16; ++r) hA[r][n] = acc[r];
  }
  __syncthreads();
  ln_relu_drop_rows(hA, P.ln1w, P.ln1b, row0, n_rows, P.T, AVHOT_SITE_EARLY0, P.ds);
  __syncthreads();
  {
    float acc[HRT16];
    const float b = __ldg(P.b4 + n);
#pragma unroll
    for (int r = 0; r < HRT16; ++r) acc[r] = b;
    const float* w = P.w4t + n;
#pragma unroll 2
    for (int k = 0; k < HD; k += 4) {
      float c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c[u] = __ldg(w + (size_t)(k + u) * HD);
#pragma unroll
      for (int r = 0; r < HRT16; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(&hA[r][k]);
        acc[r] = fmaf(c[0], x.x, acc[r]); acc[r] = fmaf(c[1], x.y, acc[r]);
        acc[r] = fmaf(c[2], x.z, acc[r]); acc[r] = fmaf(c[3], x.w, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < HRT16; ++r) hB[r][n] = acc[r];
  }
  __syncthreads();
  ln_relu_drop_rows(hB, P.ln5w, P.ln5b, row0, n_rows, P.T, AVHOT_SITE_EARLY1, P.ds);
  __syncthreads();
  const int r = n >> 4, c = n & 15;
  if (c < P.C && row0 + r < n_rows) {
    const float4* wc = reinterpret_cast<const float4*>(P.w8 + (size_t)c * HD);
    const float4* x = reinterpret_cast<const float4*>(&hB[r][0]);
    float a = 0.f;
#pragma unroll 8
    for (int k = 0; k < HD / 4; ++k) {
      const float4 u = __ldg(wc + k), v = x[k];
      a = fmaf(u.x, v.x, a); a = fmaf(u.y, v.y, a); a = fmaf(u.z, v.z, a); a = fmaf(u.w, v.w, a);
    }
    P.logits[(size_t)(row0 + r) * P.C + c] = a + __ldg(P.b8 + c);
  }
}

struct LateParams {
  const float* z[AVHOT_MAX_MODALITIES];
  int rpc[AVHOT_MAX_MODALITIES], dim[AVHOT_MAX_MODALITIES];
  const float* mask;
  float *logits, *per_modality;        // per_modality: (M, R, C) or NULL
  int M, B, T, C, present;             // present: bit m set = modality m was provided
  const float *w0t[AVHOT_MAX_MODALITIES], *b0[AVHOT_MAX_MODALITIES], *w3[AVHOT_MAX_MODALITIES], *b3[AVHOT_MAX_MODALITIES];
  const float* fusion_logits;
  DropSpec ds;
};

__global__ void __launch_bounds__(HTH) late_fusion_kernel(const LateParams P) {
  __shared__ __align__(16) float hid[HRT16][HD];
  __shared__ float wgt[HRT16][AVHOT_MAX_MODALITIES];
  const int n = threadIdx.x;
  const long long n_rows = (long long)P.B * P.T, row0 = (long long)blockIdx.x * HRT16;
  if (n < HRT16) {     // fusion weights of row n: softmax(fusion_logits) * mask, renormalised (src/fusion.py:236-243)
    const HeadRow hr = head_row(row0, n, n_rows, P.T);
    float mx = -INFINITY, e[AVHOT_MAX_MODALITIES], den = 0.f;
    for (int m = 0; m < P.M; ++m) mx = fmaxf(mx, __ldg(P.fusion_logits + m));
    for (int m = 0; m < P.M; ++m) { e[m] = expf(__ldg(P.fusion_logits + m) - mx); den += e[m]; }
    float sum = 0.f;
    for (int m = 0; m < P.M; ++m) {
      e[m] /= den;
      if (P.mask) e[m] *= __ldg(P.mask + (size_t)hr.b * P.M + m);
      sum += e[m];
    }
    for (int m = 0; m < P.M; ++m) wgt[n][m] = P.mask ? e[m] / fmaxf(sum, 1e-8f) : e[m];
  }
  const int r2 = n >> 4, c2 = n & 15;
  float fused = 0.f;
  for (int m = 0; m < P.M; ++m) {
    __syncthreads();                                  // hid reuse; wgt visible
    const bool present = (P.present >> m) & 1;
    if (present) {
      float acc[HRT16];
      const float b = __ldg(P.b0[m] + n);
      const float* src[HRT16];
#pragma unroll
      for (int r = 0; r < HRT16; ++r) {
        acc[r] = b;
        const HeadRow hr = head_row(row0, r, n_rows, P.T);
        src[r] = P.z[m] + (size_t)((long long)hr.b * P.rpc[m] + (P.rpc[m] == 1 ? 0 : hr.t)) * P.dim[m];
      }
      const float* w = P.w0t[m] + n;
      for (int k = 0; k < P.dim[m]; k += 4) {
        float c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = __ldg(w + (size_t)(k + u) * HD);
#pragma unroll
        for (int r = 0; r < HRT16; ++r) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(src[r] + k));
          acc[r] = fmaf(c[0], x.x, acc[r]); acc[r] = fmaf(c[1], x.y, acc[r]);
          acc[r] = fmaf(c[2], x.z, acc[r]); acc[r] = fmaf(c[3], x.w, acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < HRT16; ++r) {
        float v = fmaxf(acc[r], 0.f);
        if (P.ds.thr != 0u) {                          // site LATE + m, element n
          const HeadRow hr = head_row(row0, r, n_rows, P.T);
          const Philox4 pw = philox4x32_10((uint32_t)(n >> 2), (uint32_t)(P.ds.clip0 + (uint64_t)hr.b),
                                           P.ds.t0 + (uint32_t)hr.t, AVHOT_SITE_LATE + m, P.ds.k0, P.ds.k1);
          v = philox_word(pw, n & 3) >= P.ds.thr ? v * P.ds.scale : 0.f;
        }
        hid[r][n] = v;
      }
    }
    __syncthreads();
    if (c2 < P.C) {
      float a = 0.f;                                   // a missing modality contributes zero logits (:222-225)
      if (present) {
        const float4* wc = reinterpret_cast<const float4*>(P.w3[m] + (size_t)c2 * HD);
        const float4* x = reinterpret_cast<const float4*>(&hid[r2][0]);
#pragma unroll 8
        for (int k = 0; k < HD / 4; ++k) {
          const float4 u = __ldg(wc + k), v = x[k];
          a = fmaf(u.x, v.x, a); a = fmaf(u.y, v.y, a); a = fmaf(u.z, v.z, a); a = fmaf(u.w, v.w, a);
        }
        a += __ldg(P.b3[m] + c2);
      }
      fused = fmaf(wgt[r2][m], a, fused);
      if (P.per_modality && row0 + r2 < n_rows) P.per_modality[((size_t)m * n_rows + row0 + r2) * P.C + c2] = a;
    }
  }
  if (c2 < P.C && row0 + r2 < n_rows) P.logits[(size_t)(row0 + r2) * P.C + c2] = fused;
}

// ---- calibration: per-sample confidence / prediction / NLL from logits, binned ---------------------------------------
// acc layout (doubles): [0, nb) bin counts, [nb, 2nb) sum of confidences, [2nb, 3nb) correct counts,
// [3nb] sum of NLL, [3nb + 1] correct total, [3nb + 2] sample count
constexpr int CAL_MAX_BINS = 64;
__global__ void __launch_bounds__(256)
calibration_kernel(const float* __restrict__ logits, const float* __restrict__ conf_in, const long long* __restrict__ pred_in,
                   const long long* __restrict__ labels, long long N, int C, const float* __restrict__ edges, int nb,
                   double* __restrict__ acc) {
  __shared__ double s_cnt[CAL_MAX_BINS], s_conf[CAL_MAX_BINS], s_cor[CAL_MAX_BINS], s_nll, s_cor_tot, s_n;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) { s_cnt[i] = 0; s_conf[i] = 0; s_cor[i] = 0; }
  if (threadIdx.x == 0) { s_nll = 0; s_cor_tot = 0; s_n = 0; }
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    float conf;
    long long pred;
    const long long y = labels[i];
    if (logits) {
      const float* lg = logits + (size_t)i * C;
      float mx = lg[0];
      int arg = 0;
      for (int c = 1; c < C; ++c) if (lg[c] > mx) { mx = lg[c]; arg = c; }        // first maximum (torch.max)
      float den = 0.f;
      for (int c = 0; c < C; ++c) den += expf(lg[c] - mx);
      conf = 1.f / den;                                                           // softmax probability of the argmax
      pred = arg;
      const float ly = (y >= 0 && y < C) ? lg[y] : mx;
      atomicAdd(&s_nll, (double)(logf(den) + mx - ly));                           // cross-entropy, reduction='sum'
    } else {
      conf = conf_in[i];
      pred = pred_in[i];
    }
    conf = fminf(fmaxf(conf, 0.f), 1.f);                                          // .clamp(0, 1)
    int bin = -1;
    for (int b = 0; b < nb; ++b) {
      const bool in = conf >= edges[b] && (b < nb - 1 ? conf < edges[b + 1] : conf <= edges[b + 1]);
      if (in) { bin = b; break; }
    }
    const double cor = pred == y ? 1.0 : 0.0;
    if (bin >= 0) {
      atomicAdd(&s_cnt[bin], 1.0);
      atomicAdd(&s_conf[bin], (double)conf);
      atomicAdd(&s_cor[bin], cor);
    }
    atomicAdd(&s_cor_tot, cor);
    atomicAdd(&s_n, 1.0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    if (s_cnt[i] != 0) {
      atomicAdd(acc + i, s_cnt[i]);
      atomicAdd(acc + nb + i, s_conf[i]);
      atomicAdd(acc + 2 * nb + i, s_cor[i]);
    }
  }
  if (threadIdx.x == 0) {
    atomicAdd(acc + 3 * nb, s_nll);
    atomicAdd(acc + 3 * nb + 1, s_cor_tot);
    atomicAdd(acc + 3 * nb + 2, s_n);
  }
}

// ---- LateFusionWithUncertainty (src/fusion.py:589-655) + UncertaintyWeightedFusion (src/uncertainty.py:279-318) ------
// inverse-uncertainty weights of one row: w_m = (mask_m / (u_m + eps)) / (sum + eps)
__device__ __forceinline__ void unc_weights(const float* u, const float* mask, int M, float eps, float* w) {
  float sum = 0.f;
  for (int m = 0; m < M; ++m) { w[m] = (1.f / (u[m] + eps)) * mask[m]; sum += w[m]; }
  for (int m = 0; m < M; ++m) w[m] = w[m] / (sum + eps);
}
__device__ __forceinline__ float softplus_ref(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // F.softplus defaults

struct UncParams {
  const float* z[AVHOT_MAX_MODALITIES];
  int rpc[AVHOT_MAX_MODALITIES], dim[AVHOT_MAX_MODALITIES];
  const float* mask;
  float *logits, *per_modality, *weights;     // (R, C); (R, M, C) or NULL; (R, M) or NULL
  int M, B, T, C, D, present;
  const float *w0t[2][AVHOT_MAX_MODALITIES], *b0[2][AVHOT_MAX_MODALITIES], *w3[2][AVHOT_MAX_MODALITIES], *b3[2][AVHOT_MAX_MODALITIES];
  float eps;
  DropSpec ds;
};

constexpr int UNC_DMAX = 256;   // largest modality feature dim staged in shared memory

__global__ void __launch_bounds__(HTH) unc_fusion_kernel(const UncParams P) {
  __shared__ __align__(16) float xs[HRT16][UNC_DMAX];
  __shared__ __align__(16) float hid[HRT16][HD];
  __shared__ float lg[HRT16][AVHOT_MAX_MODALITIES][HCMAX];
  __shared__ float us[HRT16][AVHOT_MAX_MODALITIES];
  const int n = threadIdx.x, lane = n & 31, warp = n >> 5;
  const long long n_rows = (long long)P.B * P.T, row0 = (long long)blockIdx.x * HRT16;
  const int r2 = n >> 4, c2 = n & 15;
  for (int m = 0; m < P.M; ++m) {
    const int d = P.dim[m];
    const bool present = (P.present >> m) & 1;
    for (int head = 0; head < 2; ++head) {                    // 0: classifier, 1: uncertainty head
      const uint32_t site_in = AVHOT_SITE_UNC + 4u * m + 2u * head, site_hid = site_in + 1u;
      __syncthreads();
      // input dropout (classifiers[m][0] / uncert_heads[m][0]); a modality that is not given is zeros (:629)
      for (int idx = n; idx < HRT16 * d; idx += HTH) {
        const int r = idx / d, k = idx % d;
        const HeadRow hr = head_row(row0, r, n_rows, P.T);
        float x = present ? __ldg(P.z[m] + (size_t)((long long)hr.b * P.rpc[m] + (P.rpc[m] == 1 ? 0 : hr.t)) * d + k) : 0.f;
        if (P.ds.thr != 0u) {
          const Philox4 pw = philox4x32_10((uint32_t)(k >> 2), (uint32_t)(P.ds.clip0 + (uint64_t)hr.b),
                                           P.ds.t0 + (uint32_t)hr.t, site_in, P.ds.k0, P.ds.k1);
          x = philox_word(pw, k & 3) >= P.ds.thr ? x * P.ds.scale : 0.f;
        }
        xs[r][k] = x;
      }
      __syncthreads();
      const float* src;                                         // the rows the output Linear reads
      int kdim;
      if (P.D > 0) {
        float acc[HRT16];
        const float b = __ldg(P.b0[head][m] + n);
#pragma unroll
        for (int r = 0; r < HRT16; ++r) acc[r] = b;
        const float* w = P.w0t[head][m] + n;
        for (int k = 0; k < d; k += 4) {
          float c[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) c[u] = __ldg(w + (size_t)(k + u) * HD);
#pragma unroll
          for (int r = 0; r < HRT16; ++r) {
            const float4 x = *reinterpret_cast<const float4*>(&xs[r][k]);
            acc[r] = fmaf(c[0], x.x, acc[r]); acc[r] = fmaf(c[1], x.y, acc[r]);
            acc[r] = fmaf(c[2], x.z, acc[r]); acc[r] = fmaf(c[3], x.w, acc[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < HRT16; ++r) {
          float v = fmaxf(acc[r], 0.f);
          if (P.ds.thr != 0u) {
            const HeadRow hr = head_row(row0, r, n_rows, P.T);
            const Philox4 pw = philox4x32_10((uint32_t)(n >> 2), (uint32_t)(P.ds.clip0 + (uint64_t)hr.b),
                                             P.ds.t0 + (uint32_t)hr.t, site_hid, P.ds.k0, P.ds.k1);
            v = philox_word(pw, n & 3) >= P.ds.thr ? v * P.ds.scale : 0.f;
          }
          hid[r][n] = v;
        }
        __syncthreads();
        src = &hid[0][0]; kdim = HD;
      } else {
        src = &xs[0][0]; kdim = d;
      }
      const int pitch = P.D > 0 ? HD : UNC_DMAX;
      if (head == 0) {
        if (c2 < P.C) {
          const float* wc = P.w3[0][m] + (size_t)c2 * kdim;
          const float* x = src + (size_t)r2 * pitch;
          float a = 0.f;
          for (int k = 0; k < kdim; k += 4) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(wc + k)), v = *reinterpret_cast<const float4*>(x + k);
            a = fmaf(u.x, v.x, a); a = fmaf(u.y, v.y, a); a = fmaf(u.z, v.z, a); a = fmaf(u.w, v.w, a);
          }
          lg[r2][m][c2] = a + __ldg(P.b3[0][m] + c2);
        }
      } else {
        for (int r = warp; r < HRT16; r += HTH / 32) {          // warp = row: dot product with the 1-row weight
          const float* x = src + (size_t)r * pitch;
          float a = 0.f;
          for (int k = lane; k < kdim; k += 32) a = fmaf(__ldg(P.w3[1][m] + k), x[k], a);
          a = warp_sum(a);
          if (lane == 0) us[r][m] = softplus_ref(a + __ldg(P.b3[1][m]));
        }
      }
    }
  }
  __syncthreads();
  if (c2 < P.C && row0 + r2 < n_rows) {
    const HeadRow hr = head_row(row0, r2, n_rows, P.T);
    float mk[AVHOT_MAX_MODALITIES], w[AVHOT_MAX_MODALITIES];
    for (int m = 0; m < P.M; ++m) mk[m] = __ldg(P.mask + (size_t)hr.b * P.M + m);
    unc_weights(us[r2], mk, P.M, P.eps, w);
    float fused = 0.f;
    for (int m = 0; m < P.M; ++m) {
      fused = fmaf(w[m], lg[r2][m][c2], fused);
      if (P.per_modality) P.per_modality[((size_t)(row0 + r2) * P.M + m) * P.C + c2] = lg[r2][m][c2];
    }
    P.logits[(size_t)(row0 + r2) * P.C + c2] = fused;
    if (P.weights && c2 == 0)
      for (int m = 0; m < P.M; ++m) P.weights[(size_t)(row0 + r2) * P.M + m] = w[m];
  }
}

// UncertaintyWeightedFusion.forward on its own: logits (R, M, C), uncert (R, M), mask (R, M) -> fused (R, C), weights (R, M)
__global__ void unc_fuse_kernel(const float* __restrict__ logits, const float* __restrict__ uncert,
                                const float* __restrict__ mask, float* __restrict__ fused, float* __restrict__ weights,
                                long long R, int M, int C, float eps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * C) return;
  const long long r = i / C;
  const int c = (int)(i % C);
  float u[AVHOT_MAX_MODALITIES], mk[AVHOT_MAX_MODALITIES], w[AVHOT_MAX_MODALITIES];
  for (int m = 0; m < M; ++m) { u[m] = uncert[r * M + m]; mk[m] = mask[r * M + m]; }
  unc_weights(u, mk, M, eps, w);
  float a = 0.f;
  for (int m = 0; m < M; ++m) a = fmaf(w[m], logits[(r * M + m) * C + c], a);
  fused[i] = a;
  if (c == 0)
    for (int m = 0; m < M; ++m) weights[r * M + m] = w[m];
}

static size_t alh(size_t n_floats) { return (n_floats * sizeof(float) + 255) & ~(size_t)255; }

}  // namespace avhot

using namespace avhot;

static bool copy_to(float* dst, const float* src, size_t n, cudaStream_t s) {
  return src && cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s) == cudaSuccess;
}
static bool transpose_to(float* dst, const float* src, int N, int K, cudaStream_t s) {
  if (!src) return false;
  transpose_rows_kernel<<<N, 128, 0, s>>>(src, dst, N, K);
  return cudaGetLastError() == cudaSuccess;
}

extern "C" int avhot_early_create(const avhot_early_desc* d, avhot_head_t* out, avhot_stream_t stream) {
  AVHOT_REQUIRE(d && out, AVHOT_ERR_INVALID, "early_create: null argument");
  AVHOT_REQUIRE(d->num_modalities >= 1 && d->num_modalities <= AVHOT_MAX_MODALITIES && d->hidden_dim == HD &&
                    d->num_classes >= 1 && d->num_classes <= HCMAX, AVHOT_ERR_UNSUPPORTED,
                "early_create: this build needs 1..4 modalities, hidden_dim 256, <= 16 classes (got M=%d D=%d C=%d)",
                d->num_modalities, d->hidden_dim, d->num_classes);
  int in_dim = 0;
  for (int m = 0; m < d->num_modalities; ++m) {
    AVHOT_REQUIRE(d->modality_dim[m] > 0 && (d->modality_dim[m] & 3) == 0, AVHOT_ERR_UNSUPPORTED,
                  "early_create: modality_dim[%d]=%d must be a positive multiple of 4", m, d->modality_dim[m]);
    in_dim += d->modality_dim[m];
  }
  AVHOT_REQUIRE(in_dim <= 2048, AVHOT_ERR_UNSUPPORTED, "early_create: concatenated dim %d > 2048", in_dim);
  cudaStream_t s = as_stream(stream);
  avhot_head* h = new avhot_head();
  memset(h, 0, sizeof(*h));
  h->kind = 0; h->M = d->num_modalities; h->D = HD; h->C = d->num_classes; h->in_dim = in_dim;
  for (int m = 0; m < h->M; ++m) h->dim[m] = d->modality_dim[m];
  const size_t total = alh((size_t)in_dim * HD) + alh((size_t)HD * HD) + 6 * alh(HD) + alh((size_t)h->C * HD) + alh(h->C);
  if (cudaMalloc(&h->arena, total) != cudaSuccess) {
    delete h;
    set_error("early_create: cudaMalloc(%zu) failed", total);
    return AVHOT_ERR_CUDA;
  }
  char* p = reinterpret_cast<char*>(h->arena);
  auto take = [&](size_t n) { float* r = reinterpret_cast<float*>(p); p += alh(n); return r; };
  h->w0t = take((size_t)in_dim * HD); h->w4t = take((size_t)HD * HD);
  h->b0 = take(HD); h->ln1w = take(HD); h->ln1b = take(HD); h->b4 = take(HD); h->ln5w = take(HD); h->ln5b = take(HD);
  h->w8 = take((size_t)h->C * HD); h->b8 = take(h->C);
  bool ok = transpose_to(h->w0t, d->w0, HD, in_dim, s) && transpose_to(h->w4t, d->w4, HD, HD, s) &&
            copy_to(h->b0, d->b0, HD, s) && copy_to(h->ln1w, d->ln1_w, HD, s) && copy_to(h->ln1b, d->ln1_b, HD, s) &&
            copy_to(h->b4, d->b4, HD, s) && copy_to(h->ln5w, d->ln5_w, HD, s) && copy_to(h->ln5b, d->ln5_b, HD, s) &&
            copy_to(h->w8, d->w8, (size_t)h->C * HD, s) && copy_to(h->b8, d->b8, h->C, s);
  if (!ok) {
    cudaFree(h->arena);
    delete h;
    set_error("early_create: missing tensor or device copy failure");
    return AVHOT_ERR_INVALID;
  }
  *out = h;
  return AVHOT_OK;
}

extern "C" int avhot_late_create(const avhot_late_desc* d, avhot_head_t* out, avhot_stream_t stream) {
  AVHOT_REQUIRE(d && out, AVHOT_ERR_INVALID, "late_create: null argument");
  AVHOT_REQUIRE(d->num_modalities >= 1 && d->num_modalities <= AVHOT_MAX_MODALITIES && d->hidden_dim == HD &&
                    d->num_classes >= 1 && d->num_classes <= HCMAX, AVHOT_ERR_UNSUPPORTED,
                "late_create: this build needs 1..4 modalities, hidden_dim 256, <= 16 classes (got M=%d D=%d C=%d)",
                d->num_modalities, d->hidden_dim, d->num_classes);
  cudaStream_t s = as_stream(stream);
  avhot_head* h = new avhot_head();
  memset(h, 0, sizeof(*h));
  h->kind = 1; h->M = d->num_modalities; h->D = HD; h->C = d->num_classes;
  size_t total = alh(AVHOT_MAX_MODALITIES);
  for (int m = 0; m < h->M; ++m) {
    if (d->modality_dim[m] <= 0 || (d->modality_dim[m] & 3)) {
      delete h;
      set_error("late_create: modality_dim[%d]=%d must be a positive multiple of 4", m, d->modality_dim[m]);
      return AVHOT_ERR_UNSUPPORTED;
    }
    h->dim[m] = d->modality_dim[m];
    total += alh((size_t)h->dim[m] * HD) + alh(HD) + alh((size_t)h->C * HD) + alh(h->C);
  }
  if (cudaMalloc(&h->arena, total) != cudaSuccess) {
    delete h;
    set_error("late_create: cudaMalloc(%zu) failed", total);
    return AVHOT_ERR_CUDA;
  }
  char* p = reinterpret_cast<char*>(h->arena);
  auto take = [&](size_t n) { float* r = reinterpret_cast<float*>(p); p += alh(n); return r; };
  bool ok = true;
  for (int m = 0; m < h->M; ++m) {
    h->lw0t[m] = take((size_t)h->dim[m] * HD); h->lb0[m] = take(HD);
    h->lw3[m] = take((size_t)h->C * HD); h->lb3[m] = take(h->C);
    ok = ok && transpose_to(h->lw0t[m], d->w0[m], HD, h->dim[m], s) && copy_to(h->lb0[m], d->b0[m], HD, s) &&
         copy_to(h->lw3[m], d->w3[m], (size_t)h->C * HD, s) && copy_to(h->lb3[m], d->b3[m], h->C, s);
  }
  h->fusion_logits = take(AVHOT_MAX_MODALITIES);
  ok = ok && copy_to(h->fusion_logits, d->fusion_logits, h->M, s);
  if (!ok) {
    cudaFree(h->arena);
    delete h;
    set_error("late_create: missing tensor or device copy failure");
    return AVHOT_ERR_INVALID;
  }
  *out = h;
  return AVHOT_OK;
}

extern "C" int avhot_unc_create(const avhot_unc_desc* d, avhot_head_t* out, avhot_stream_t stream) {
  AVHOT_REQUIRE(d && out, AVHOT_ERR_INVALID, "unc_create: null argument");
  AVHOT_REQUIRE(d->num_modalities >= 1 && d->num_modalities <= AVHOT_MAX_MODALITIES && (d->hidden_dim == HD || d->hidden_dim == 0) &&
                    d->num_classes >= 1 && d->num_classes <= HCMAX, AVHOT_ERR_UNSUPPORTED,
                "unc_create: this build needs 1..4 modalities, hidden_dim 0 or 256, <= 16 classes (got M=%d D=%d C=%d)",
                d->num_modalities, d->hidden_dim, d->num_classes);
  cudaStream_t s = as_stream(stream);
  avhot_head* h = new avhot_head();
  memset(h, 0, sizeof(*h));
  h->kind = 2; h->M = d->num_modalities; h->D = d->hidden_dim; h->C = d->num_classes; h->eps = d->epsilon;
  const int D = h->D;
  size_t total = 0;
  for (int m = 0; m < h->M; ++m) {
    const int dm = d->modality_dim[m];
    if (dm <= 0 || (dm & 3) || dm > UNC_DMAX) {
      delete h;
      set_error("unc_create: modality_dim[%d]=%d must be a positive multiple of 4, <= %d", m, dm, UNC_DMAX);
      return AVHOT_ERR_UNSUPPORTED;
    }
    h->dim[m] = dm;
    for (int head = 0; head < 2; ++head) {
      const int outs = head == 0 ? h->C : 1;
      total += D > 0 ? alh((size_t)dm * D) + alh(D) + alh((size_t)outs * D) + alh(outs) : alh((size_t)outs * dm) + alh(outs);
    }
  }
  if (cudaMalloc(&h->arena, total) != cudaSuccess) {
    delete h;
    set_error("unc_create: cudaMalloc(%zu) failed", total);
    return AVHOT_ERR_CUDA;
  }
  char* p = reinterpret_cast<char*>(h->arena);
  auto take = [&](size_t n) { float* r = reinterpret_cast<float*>(p); p += alh(n); return r; };
  bool ok = true;
  for (int m = 0; m < h->M && ok; ++m)
    for (int head = 0; head < 2 && ok; ++head) {
      const int outs = head == 0 ? h->C : 1, dm = h->dim[m];
      const float* w_in = head == 0 ? d->cls_w0[m] : d->unc_w0[m];
      const float* b_in = head == 0 ? d->cls_b0[m] : d->unc_b0[m];
      const float* w_out = head == 0 ? d->cls_w3[m] : d->unc_w3[m];
      const float* b_out = head == 0 ? d->cls_b3[m] : d->unc_b3[m];
      if (D > 0) {
        h->uw0t[head][m] = take((size_t)dm * D); h->ub0[head][m] = take(D);
        h->uw3[head][m] = take((size_t)outs * D); h->ub3[head][m] = take(outs);
        ok = transpose_to(h->uw0t[head][m], w_in, D, dm, s) && copy_to(h->ub0[head][m], b_in, D, s) &&
             copy_to(h->uw3[head][m], w_out, (size_t)outs * D, s) && copy_to(h->ub3[head][m], b_out, outs, s);
      } else {                                                  // a single Linear(d -> outs): w3 / b3 slots
        h->uw3[head][m] = take((size_t)outs * dm); h->ub3[head][m] = take(outs);
        ok = copy_to(h->uw3[head][m], w_out, (size_t)outs * dm, s) && copy_to(h->ub3[head][m], b_out, outs, s);
      }
    }
  if (!ok) {
    cudaFree(h->arena);
    delete h;
    set_error("unc_create: missing tensor or device copy failure");
    return AVHOT_ERR_INVALID;
  }
  *out = h;
  return AVHOT_OK;
}

extern "C" int avhot_unc_forward(avhot_head_t h, const float* const* z_host_array, const int* z_rows_per_clip_host,
                                 int present_mask, const float* mask, float* logits, float* per_modality_logits,
                                 float* fusion_weights, int B, int T, float p, uint64_t seed, uint64_t clip0, uint32_t t0,
                                 avhot_stream_t stream) {
  AVHOT_REQUIRE(h && h->kind == 2 && z_host_array && z_rows_per_clip_host && logits && mask && B >= 0 && T >= 1,
                AVHOT_ERR_INVALID, "unc_forward: bad arguments (modality_mask is required, src/fusion.py:622)");
  AVHOT_REQUIRE(p >= 0.f && p < 1.f, AVHOT_ERR_INVALID, "unc_forward: dropout p=%g outside [0, 1)", (double)p);
  if (B == 0) return AVHOT_OK;
  const long long rows = (long long)B * T, grid = (rows + HRT16 - 1) / HRT16;
  AVHOT_REQUIRE(grid <= 0x7fffffffLL, AVHOT_ERR_INVALID, "unc_forward: B*T too large for one call");
  UncParams P;
  memset(&P, 0, sizeof(P));
  for (int m = 0; m < h->M; ++m) {
    const bool present = (present_mask >> m) & 1;
    AVHOT_REQUIRE(!present || (z_host_array[m] && (z_rows_per_clip_host[m] == 1 || z_rows_per_clip_host[m] == T)),
                  AVHOT_ERR_INVALID, "unc_forward: z[%d] is NULL or z_rows_per_clip is not 1 / T", m);
    P.z[m] = z_host_array[m]; P.rpc[m] = present ? z_rows_per_clip_host[m] : 1; P.dim[m] = h->dim[m];
    for (int head = 0; head < 2; ++head) {
      P.w0t[head][m] = h->uw0t[head][m]; P.b0[head][m] = h->ub0[head][m];
      P.w3[head][m] = h->uw3[head][m]; P.b3[head][m] = h->ub3[head][m];
    }
  }
  P.mask = mask; P.logits = logits; P.per_modality = per_modality_logits; P.weights = fusion_weights;
  P.M = h->M; P.B = B; P.T = T; P.C = h->C; P.D = h->D; P.present = present_mask; P.eps = h->eps;
  P.ds = make_drop(p, seed, clip0, t0);
  ProfScope prof("uncertainty_fusion", as_stream(stream));
  unc_fusion_kernel<<<(unsigned)grid, HTH, 0, as_stream(stream)>>>(P);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

extern "C" int avhot_uncertainty_fuse(const float* logits, const float* uncert, const float* mask, float* fused,
                                      float* weights, long long R, int M, int C, float eps, avhot_stream_t stream) {
  AVHOT_REQUIRE(logits && uncert && mask && fused && weights && R >= 0 && M >= 1 && M <= AVHOT_MAX_MODALITIES && C >= 1,
                AVHOT_ERR_INVALID, "uncertainty_fuse: bad arguments (1..4 modalities)");
  if (R == 0) return AVHOT_OK;
  const long long n = R * C;
  AVHOT_REQUIRE((n + 255) / 256 <= 0x7fffffffLL, AVHOT_ERR_INVALID, "uncertainty_fuse: too many rows");
  ProfScope prof("uncertainty_fuse", as_stream(stream));
  unc_fuse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(logits, uncert, mask, fused, weights, R, M, C, eps);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

extern "C" void avhot_head_destroy(avhot_head_t h) {
  if (!h) return;
  cudaFree(h->arena);
  delete h;
}

extern "C" int avhot_head_forward(avhot_head_t h, const float* const* z_host_array, const int* z_rows_per_clip_host,
                                  int present_mask, const float* mask, float* logits, float* per_modality_logits,
                                  int B, int T, float p, uint64_t seed, uint64_t clip0, uint32_t t0,
                                  avhot_stream_t stream) {
  AVHOT_REQUIRE(h && z_host_array && z_rows_per_clip_host && logits && B >= 0 && T >= 1, AVHOT_ERR_INVALID,
                "head_forward: bad arguments");
  AVHOT_REQUIRE(p >= 0.f && p < 1.f, AVHOT_ERR_INVALID, "head_forward: dropout p=%g outside [0, 1)", (double)p);
  if (B == 0) return AVHOT_OK;
  const long long rows = (long long)B * T, grid = (rows + HRT16 - 1) / HRT16;
  AVHOT_REQUIRE(grid <= 0x7fffffffLL, AVHOT_ERR_INVALID, "head_forward: B*T too large for one call");
  for (int m = 0; m < h->M; ++m) {
    const bool present = (present_mask >> m) & 1;
    AVHOT_REQUIRE(!present || z_host_array[m], AVHOT_ERR_INVALID, "head_forward: z[%d] is NULL", m);
    AVHOT_REQUIRE(!present || z_rows_per_clip_host[m] == 1 || z_rows_per_clip_host[m] == T, AVHOT_ERR_INVALID,
                  "head_forward: z_rows_per_clip[%d] must be 1 or T", m);
    AVHOT_REQUIRE(!present || (reinterpret_cast<uintptr_t>(z_host_array[m]) & 15) == 0, AVHOT_ERR_INVALID,
                  "head_forward: z[%d] must be 16-byte aligned", m);
  }
  cudaStream_t s = as_stream(stream);
  if (h->kind == 0) {
    AVHOT_REQUIRE(present_mask == (1 << h->M) - 1, AVHOT_ERR_INVALID,
                  "head_forward(early): pass zeros for a missing modality (src/fusion.py:120-121)");
    AVHOT_REQUIRE(!per_modality_logits, AVHOT_ERR_INVALID, "head_forward(early): no per-modality logits");
    EarlyParams P;
    for (int m = 0; m < h->M; ++m) { P.z[m] = z_host_array[m]; P.rpc[m] = z_rows_per_clip_host[m]; P.dim[m] = h->dim[m]; }
    P.mask = mask; P.logits = logits; P.M = h->M; P.B = B; P.T = T; P.C = h->C; P.in_dim = h->in_dim;
    P.w0t = h->w0t; P.b0 = h->b0; P.ln1w = h->ln1w; P.ln1b = h->ln1b; P.w4t = h->w4t; P.b4 = h->b4;
    P.ln5w = h->ln5w; P.ln5b = h->ln5b; P.w8 = h->w8; P.b8 = h->b8;
    P.ds = make_drop(p, seed, clip0, t0);
    const size_t smem = sizeof(float) * HRT16 * (size_t)(h->in_dim + 2 * HD);
    AVHOT_CUDA(cudaFuncSetAttribute(early_fusion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof("early_fusion", s);
    early_fusion_kernel<<<(unsigned)grid, HTH, smem, s>>>(P);
    AVHOT_LAUNCH_CHECK();
  } else {
    LateParams P;
    for (int m = 0; m < h->M; ++m) {
      P.z[m] = z_host_array[m]; P.rpc[m] = z_rows_per_clip_host[m]; P.dim[m] = h->dim[m];
      P.w0t[m] = h->lw0t[m]; P.b0[m] = h->lb0[m]; P.w3[m] = h->lw3[m]; P.b3[m] = h->lb3[m];
    }
    P.mask = mask; P.logits = logits; P.per_modality = per_modality_logits;
    P.M = h->M; P.B = B; P.T = T; P.C = h->C; P.present = present_mask; P.fusion_logits = h->fusion_logits;
    P.ds = make_drop(p, seed, clip0, t0);
    ProfScope prof("late_fusion", s);
    late_fusion_kernel<<<(unsigned)grid, HTH, 0, s>>>(P);
    AVHOT_LAUNCH_CHECK();
  }
  return AVHOT_OK;
}

extern "C" int avhot_calibration_accumulate(const float* logits, const float* confidences, const int64_t* predictions,
                                            const int64_t* labels, long long N, int C, const float* bin_edges,
                                            int num_bins, double* acc, avhot_stream_t stream) {
  AVHOT_REQUIRE(labels && bin_edges && acc && N >= 0, AVHOT_ERR_INVALID, "calibration: null argument");
  AVHOT_REQUIRE((logits != nullptr) != (confidences != nullptr && predictions != nullptr), AVHOT_ERR_INVALID,
                "calibration: give either logits or (confidences, predictions)");
  AVHOT_REQUIRE(num_bins >= 1 && num_bins <= CAL_MAX_BINS && (!logits || (C >= 1 && C <= 4096)), AVHOT_ERR_INVALID,
                "calibration: 1 <= num_bins <= 64");
  if (N == 0) return AVHOT_OK;
  const int grid = (int)(((N + 255) / 256) < (long long)(4 * num_sms()) ? ((N + 255) / 256) : (long long)(4 * num_sms()));
  ProfScope prof("calibration", as_stream(stream));
  calibration_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, confidences, reinterpret_cast<const long long*>(predictions),
                                                        reinterpret_cast<const long long*>(labels), N, C, bin_edges,
                                                        num_bins, acc);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}
