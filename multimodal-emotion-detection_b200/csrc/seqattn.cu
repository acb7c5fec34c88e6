// the sequence-level attention modules of the reference.
//
//   SequenceEncoder(encoder_type='transformer')   src/encoders.py:96-113 (ctor), :170-192 (forward)
//   TemporalAttention (+ pool_sequence)           src/attention.py:143-274
//   PairwiseModalityAttention's aggregation       src/attention.py:421-548 (the second definition, which is the one
//                                                 Python keeps); its CrossModalAttention blocks are csrc/xattn.cu
//
// fp32 parity path: every nn.Linear is the fp32 GEMM of gemm.cu; multi-head self-attention is one kernel per
// (clip, head, 32-query tile) that stages K / V tiles in shared memory and makes two sweeps over the keys (row max
// and sum, then normalised weights and P V) so the attention weights it can return are the softmax the reference
// returns.  Activations stay batch-first (B, T, H) throughout - the reference's (T, B, H) transposes
// (src/encoders.py:183,185) only change the order in which independent clips are stored.
//
// nn.TransformerEncoderLayer as the reference builds it (post-norm, GELU, batch_first=False, so neither the nested
// tensor nor the fused fast path is taken):
//     x = LN1(x + dropout1(SA(x)));  x = LN2(x + dropout2(linear2(dropout(gelu(linear1(x))))))
// Under MCDropoutUncertainty only nn.Dropout modules are stochastic (src/uncertainty.py:36-46); the attention-weight
// dropout inside nn.MultiheadAttention follows that module's own eval flag and stays off.  Sites, per layer l:
// AVHOT_SITE_TFM + 3 l + {0: dropout1 (T, H), 1: dropout (T, F), 2: dropout2 (T, H)}, element = row-major index inside
// the per-clip tensor.  Layer 0's self-attention block sees the pass-invariant input and is computed once.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace avhot {

constexpr int SA_QT = 32;        // queries per CTA
constexpr int SA_KT = 32;        // keys per staged tile (one per lane)
constexpr int SA_WARPS = 8;
constexpr int SA_QPW = SA_QT / SA_WARPS;   // queries per warp, processed together
constexpr int SA_MAX_D = 128;    // head dim

// 8 keep bits for elements e0 .. e0 + 7 (e0 % 8 == 0) of a site whose per-clip mask has n_elems elements
// (oracle/philox.py dropout_keep: 16-bit draws for dense sites, 32-bit draws otherwise)
__device__ __forceinline__ uint32_t keep8(uint32_t e0, bool dense, uint32_t clip, uint32_t pass, uint32_t site,
                                          const DropSpec& ds) {
  uint32_t bits = 0;
  if (dense) {
    const Philox4 r = philox4x32_10(e0 >> 3, clip, pass, site, ds.k0, ds.k1);
    const uint32_t thr16 = ds.thr >> 16;
#pragma unroll
    for (int q = 0; q < 8; ++q) bits |= (philox_half(r, (uint32_t)q) >= thr16 ? 1u : 0u) << q;
  } else {
    const Philox4 r0 = philox4x32_10(e0 >> 2, clip, pass, site, ds.k0, ds.k1);
    const Philox4 r1 = philox4x32_10((e0 >> 2) + 1, clip, pass, site, ds.k0, ds.k1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      bits |= (philox_word(r0, (uint32_t)q) >= ds.thr ? 1u : 0u) << q;
      bits |= (philox_word(r1, (uint32_t)q) >= ds.thr ? 1u : 0u) << (q + 4);
    }
  }
  return bits;
}

// ---- multi-head self-attention ---------------------------------------------------------------------------------
// q / k / v element (b, s, head, i) at base[(b S + s) ld + off + head d + i].  Key j of clip b is valid iff
// (lengths == NULL || j < lengths[b]) && (valid == NULL || valid[b S + j]).  Masked keys score -inf; a row whose keys
// are all masked is NaN, as torch.softmax makes it.  weights (B, heads, S, S) and query_mask (B, S) are optional
// (TemporalAttention returns the former and multiplies the context by the latter, src/attention.py:232-236).
__global__ void __launch_bounds__(SA_WARPS * 32)
self_attn_kernel(const float* __restrict__ qkv, int ld, int off_q, int off_k, int off_v, const int32_t* __restrict__ lengths,
                 const uint8_t* __restrict__ valid, const uint8_t* __restrict__ query_mask, float* __restrict__ ctx,
                 int ldo, float* __restrict__ weights, int S, int heads, int d, float scale) {
  extern __shared__ __align__(16) float sa_smem[];
  float* Qs = sa_smem;                       // [SA_QT][d]
  float* Ks = Qs + SA_QT * d;                // [SA_KT][d + 1]
  float* Vs = Ks + SA_KT * (d + 1);          // [SA_KT][d]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * SA_QT, head = blockIdx.y, b = blockIdx.z;
  const float* base = qkv + (size_t)b * S * ld;
  const int len = lengths ? min(max(lengths[b], 0), S) : S;
  const uint8_t* vb = valid ? valid + (size_t)b * S : nullptr;
  const float NEG_INF = -INFINITY;

  for (int i = tid; i < SA_QT * d; i += blockDim.x) {
    const int r = i / d, c = i % d, s = q0 + r;
    Qs[i] = s < S ? base[(size_t)s * ld + off_q + head * d + c] : 0.f;
  }
  float m[SA_QPW], l[SA_QPW];
#pragma unroll
  for (int u = 0; u < SA_QPW; ++u) { m[u] = NEG_INF; l[u] = 0.f; }
  const float* qw = Qs + (warp * SA_QPW) * d;

  auto load_tile = [&](int k0, bool with_v) {
    __syncthreads();   // the previous tile is no longer read (also orders the Qs fill before its first use)
    for (int i = tid; i < SA_KT * d; i += blockDim.x) {
      const int r = i / d, c = i % d, s = k0 + r;
      const bool in = s < S;
      Ks[r * (d + 1) + c] = in ? base[(size_t)s * ld + off_k + head * d + c] : 0.f;
      if (with_v) Vs[i] = in ? base[(size_t)s * ld + off_v + head * d + c] : 0.f;
    }
    __syncthreads();
  };
  auto scores = [&](int k0, float (&sc)[SA_QPW]) {
    const int j = k0 + lane;
    const float* kr = Ks + lane * (d + 1);
#pragma unroll
    for (int u = 0; u < SA_QPW; ++u) sc[u] = 0.f;
    for (int i = 0; i < d; ++i) {
      const float kv = kr[i];
#pragma unroll
      for (int u = 0; u < SA_QPW; ++u) sc[u] = fmaf(qw[u * d + i], kv, sc[u]);
    }
    const bool ok = j < len && (!vb || vb[j]);
#pragma unroll
    for (int u = 0; u < SA_QPW; ++u) sc[u] = ok ? sc[u] * scale : NEG_INF;
  };

  // sweep 1: row max and sum of exp
  for (int k0 = 0; k0 < S; k0 += SA_KT) {
    load_tile(k0, false);
    float sc[SA_QPW];
    scores(k0, sc);
#pragma unroll
    for (int u = 0; u < SA_QPW; ++u) {
      const float mt = warp_max(sc[u]);
      const float mn = fmaxf(m[u], mt);
      if (mn != NEG_INF) {
        const float pe = sc[u] == NEG_INF ? 0.f : expf(sc[u] - mn);
        const float ps = warp_sum(pe);
        l[u] = (m[u] == NEG_INF ? 0.f : l[u] * expf(m[u] - mn)) + ps;
        m[u] = mn;
      }
    }
  }
  // sweep 2: normalised weights, P V
  float acc[SA_QPW][SA_MAX_D / 32];
#pragma unroll
  for (int u = 0; u < SA_QPW; ++u)
#pragma unroll
    for (int c = 0; c < SA_MAX_D / 32; ++c) acc[u][c] = 0.f;
  for (int k0 = 0; k0 < S; k0 += SA_KT) {
    load_tile(k0, true);
    float sc[SA_QPW], w[SA_QPW];
    scores(k0, sc);
#pragma unroll
    for (int u = 0; u < SA_QPW; ++u) {
      if (m[u] == NEG_INF) w[u] = NAN;                       // softmax over an all -inf row
      else w[u] = sc[u] == NEG_INF ? 0.f : expf(sc[u] - m[u]) / l[u];
      const int s = q0 + warp * SA_QPW + u, j = k0 + lane;
      if (weights && s < S && j < S) weights[(((size_t)b * heads + head) * S + s) * S + j] = w[u];
      if (j >= S) w[u] = 0.f;                                 // keys past the end of the sequence do not exist
    }
    const int nk = min(SA_KT, S - k0);
    for (int j = 0; j < nk; ++j) {
      float pj[SA_QPW];
#pragma unroll
      for (int u = 0; u < SA_QPW; ++u) pj[u] = __shfl_sync(0xffffffffu, w[u], j);
#pragma unroll
      for (int c = 0; c < SA_MAX_D / 32; ++c) {
        const int col = lane + 32 * c;
        if (col < d) {
          const float vv = Vs[j * d + col];
#pragma unroll
          for (int u = 0; u < SA_QPW; ++u) acc[u][c] = fmaf(pj[u], vv, acc[u][c]);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < SA_QPW; ++u) {
    const int s = q0 + warp * SA_QPW + u;
    if (s >= S) continue;
    const float qm = query_mask ? (query_mask[(size_t)b * S + s] ? 1.f : 0.f) : 1.f;
#pragma unroll
    for (int c = 0; c < SA_MAX_D / 32; ++c) {
      const int col = lane + 32 * c;
      if (col < d) ctx[((size_t)b * S + s) * ldo + head * d + col] = query_mask ? acc[u][c] * qm : acc[u][c];
    }
  }
}

static int launch_self_attn(const float* qkv, int ld, int off_q, int off_k, int off_v, const int32_t* lengths,
                            const uint8_t* valid, const uint8_t* query_mask, float* ctx, int ldo, float* weights, int B,
                            int S, int heads, int d, cudaStream_t s, const char* prof_name) {
  AVHOT_REQUIRE(d >= 1 && d <= SA_MAX_D, AVHOT_ERR_UNSUPPORTED, "self-attention: head dim %d outside 1..%d", d, SA_MAX_D);
  AVHOT_REQUIRE(B <= 65535 && heads <= 65535, AVHOT_ERR_INVALID, "self-attention: B or heads > 65535 per call");
  const size_t smem = sizeof(float) * ((size_t)SA_QT * d + (size_t)SA_KT * (d + 1) + (size_t)SA_KT * d);
  AVHOT_CUDA(cudaFuncSetAttribute(self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(prof_name, s);
  dim3 grid(ceil_div(S, SA_QT), heads, B);
  self_attn_kernel<<<grid, SA_WARPS * 32, smem, s>>>(qkv, ld, off_q, off_k, off_v, lengths, valid, query_mask, ctx, ldo,
                                                     weights, S, heads, d, 1.0f / sqrtf((float)d));
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

// ---- row kernels ------------------------------------------------------------------------------------------------
// x[b, t, :] += pos[min(t, max_len - 1), :]      (src/encoders.py:180-182)
__global__ void add_pos_kernel(float* __restrict__ x, const float* __restrict__ pos, long long rows, int T, int H,
                               int max_len) {
  const long long n4 = rows * (H / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / (H / 4);
    const int c4 = (int)(i % (H / 4));
    const int t = min((int)(row % T), max_len - 1);
    float4 v = reinterpret_cast<float4*>(x)[i];
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (size_t)t * H) + c4);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

// out[row] = LayerNorm(x[row] + dropout(y[row])) * w + b; warp = row (b, t); H % 8 == 0, H <= 1024.
// Mask element of (clip b, row t, column h) = t H + h of a (T, H) per-clip mask.
constexpr int LN_MAX_CHUNKS = 4;   // 8-element chunks per lane
__global__ void __launch_bounds__(256)
resid_drop_ln_kernel(const float* x, const float* __restrict__ y, const float* __restrict__ w,
                     const float* __restrict__ bias, float eps, float* out /* may alias x: a warp owns its row */, long long rows, int T, int H,
                     uint32_t site, int pass, DropSpec ds) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = (int)(row / T), t = (int)(row % T);
  const bool drop = ds.thr != 0u;
  const bool dense = (long long)T * H >= AVHOT_DENSE_SITE_ELEMS;
  const uint32_t clip = (uint32_t)(ds.clip0 + (uint64_t)b);
  float v[LN_MAX_CHUNKS][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
    const int h0 = (lane + 32 * i) * 8;
    if (h0 < H) {
      const float4 x0 = *reinterpret_cast<const float4*>(x + row * H + h0), x1 = *reinterpret_cast<const float4*>(x + row * H + h0 + 4);
      const float4 y0 = *reinterpret_cast<const float4*>(y + row * H + h0), y1 = *reinterpret_cast<const float4*>(y + row * H + h0 + 4);
      const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      const float ys[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
      uint32_t bits = 0xFFu;
      if (drop) bits = keep8((uint32_t)t * (uint32_t)H + (uint32_t)h0, dense, clip, ds.t0 + (uint32_t)pass, site, ds);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float yd = drop ? ((bits >> q) & 1u ? ys[q] * ds.scale : 0.f) : ys[q];
        v[i][q] = xs[q] + yd;
        sum += v[i][q];
      }
    }
  }
  const float mean = warp_sum(sum) / (float)H;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i)
    if ((lane + 32 * i) * 8 < H)
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float dlt = v[i][q] - mean; sq = fmaf(dlt, dlt, sq); }
  const float rstd = rsqrtf(warp_sum(sq) / (float)H + eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
    const int h0 = (lane + 32 * i) * 8;
    if (h0 < H) {
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = (v[i][q] - mean) * rstd * __ldg(w + h0 + q) + __ldg(bias + h0 + q);
      *reinterpret_cast<float4*>(out + row * H + h0) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(out + row * H + h0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// a[row, j] = dropout(gelu(a[row, j])) in place, exact (erf) GELU; thread = 8 consecutive columns; F % 8 == 0
__global__ void gelu_drop_kernel(float* __restrict__ a, long long rows, int T, int F, uint32_t site, int pass, DropSpec ds) {
  const long long n8 = rows * (F / 8);
  const bool drop = ds.thr != 0u;
  const bool dense = (long long)T * F >= AVHOT_DENSE_SITE_ELEMS;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / (F / 8);
    const int j0 = (int)(i % (F / 8)) * 8;
    const int b = (int)(row / T), t = (int)(row % T);
    float4 a0 = *reinterpret_cast<float4*>(a + row * F + j0), a1 = *reinterpret_cast<float4*>(a + row * F + j0 + 4);
    float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    uint32_t bits = 0xFFu;
    if (drop) bits = keep8((uint32_t)t * (uint32_t)F + (uint32_t)j0, dense, (uint32_t)(ds.clip0 + (uint64_t)b),
                           ds.t0 + (uint32_t)pass, site, ds);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float g = 0.5f * v[q] * (1.f + erff(v[q] * 0.70710678118654752440f));
      v[q] = drop ? ((bits >> q) & 1u ? g * ds.scale : 0.f) : g;
    }
    *reinterpret_cast<float4*>(a + row * F + j0) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(a + row * F + j0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// pooled[(b P + pass), h] = sum_{t < len_b} x[b, t, h] / max(len_b, 1)   (src/encoders.py:186-192; len = T without lengths)
__global__ void seq_mean_pool_kernel(const float* __restrict__ x, const int32_t* __restrict__ lengths,
                                     float* __restrict__ pooled, int T, int H, int P, int pass) {
  const int b = blockIdx.x;
  const int len = lengths ? min(max(lengths[b], 0), T) : T;
  const float denom = (float)max(len, 1);
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < len; ++t) s += x[((size_t)b * T + t) * H + h];
    pooled[((size_t)b * P + pass) * H + h] = s / denom;
  }
}

// TemporalAttention.pool_sequence (src/attention.py:245-274): imp[k] = mean_{head, q} w[b, head, q, k];
// imp /= (sum_k imp + 1e-9); pooled[b, :] = sum_k imp[k] seq[b, k, :].  One CTA per clip.
__global__ void __launch_bounds__(256)
tattn_pool_kernel(const float* __restrict__ seq, const float* __restrict__ weights, float* __restrict__ pooled, int S,
                  int D, int heads) {
  extern __shared__ float imp[];           // [S]
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* wb = weights + (size_t)b * heads * S * S;
  float part = 0.f;
  for (int k = tid; k < S; k += blockDim.x) {
    float tot = 0.f;
    for (int hd = 0; hd < heads; ++hd) {
      float s = 0.f;
      for (int q = 0; q < S; ++q) s += wb[((size_t)hd * S + q) * S + k];
      tot += s;
    }
    const float v = tot / (float)heads / (float)S;
    imp[k] = v;
    part += v;
  }
  part = warp_sum(part);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  float total = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) total += red[i];
  const float inv = 1.f / (total + 1e-9f);
  for (int c = tid; c < D; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < S; ++k) s = fmaf(imp[k] * inv, seq[((size_t)b * S + k) * D + c], s);
    pooled[(size_t)b * D + c] = s;
  }
}

// PairwiseModalityAttention, per target modality (src/attention.py:527-545):
//   out[b] = LayerNorm(self_feat[b] + dropout(sum_j msg_j[b])) * avail[b]        warp = row; H % 8 == 0, H <= 1024
__global__ void __launch_bounds__(256)
pair_combine_kernel(const float* __restrict__ self_feat, const float* __restrict__ m0, const float* __restrict__ m1,
                    const float* __restrict__ m2, const uint8_t* __restrict__ avail, const float* __restrict__ w,
                    const float* __restrict__ bias, float eps, float* __restrict__ out, int B, int H, uint32_t site,
                    DropSpec ds) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const bool drop = ds.thr != 0u;
  const bool dense = H >= AVHOT_DENSE_SITE_ELEMS;
  const uint32_t clip = (uint32_t)(ds.clip0 + (uint64_t)row);
  float v[LN_MAX_CHUNKS][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
    const int h0 = (lane + 32 * i) * 8;
    if (h0 < H) {
      uint32_t bits = 0xFFu;
      if (drop) bits = keep8((uint32_t)h0, dense, clip, ds.t0, site, ds);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const size_t idx = (size_t)row * H + h0 + q;
        float msg = 0.f;                                    // torch.stack(messages).sum(0): left to right
        if (m0) msg = m0[idx];
        if (m1) msg += m1[idx];
        if (m2) msg += m2[idx];
        if (drop) msg = (bits >> q) & 1u ? msg * ds.scale : 0.f;
        v[i][q] = self_feat[idx] + msg;
        sum += v[i][q];
      }
    }
  }
  const float mean = warp_sum(sum) / (float)H;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i)
    if ((lane + 32 * i) * 8 < H)
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float dlt = v[i][q] - mean; sq = fmaf(dlt, dlt, sq); }
  const float rstd = rsqrtf(warp_sum(sq) / (float)H + eps);
  const float am = avail ? (avail[row] ? 1.f : 0.f) : 1.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
    const int h0 = (lane + 32 * i) * 8;
    if (h0 < H)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float o = (v[i][q] - mean) * rstd * __ldg(w + h0 + q) + __ldg(bias + h0 + q);
        out[(size_t)row * H + h0 + q] = avail ? o * am : o;
      }
  }
}

static size_t salign(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace avhot

using namespace avhot;

// =================================================================================================================
// SequenceEncoder(encoder_type='transformer')
// =================================================================================================================
struct avhot_tfm {
  int in_dim, H, out_dim, layers, heads, F, max_len;
  float eps;
  float* arena;
  float *input_w, *input_b, *pos, *proj_w, *proj_b;
  float *in_proj_w[AVHOT_TFM_MAX_LAYERS], *in_proj_b[AVHOT_TFM_MAX_LAYERS], *out_w[AVHOT_TFM_MAX_LAYERS],
      *out_b[AVHOT_TFM_MAX_LAYERS], *lin1_w[AVHOT_TFM_MAX_LAYERS], *lin1_b[AVHOT_TFM_MAX_LAYERS],
      *lin2_w[AVHOT_TFM_MAX_LAYERS], *lin2_b[AVHOT_TFM_MAX_LAYERS], *n1_w[AVHOT_TFM_MAX_LAYERS],
      *n1_b[AVHOT_TFM_MAX_LAYERS], *n2_w[AVHOT_TFM_MAX_LAYERS], *n2_b[AVHOT_TFM_MAX_LAYERS];
};

extern "C" int avhot_tfm_create(const avhot_tfm_desc* d, avhot_tfm_t* out, avhot_stream_t stream) {
  AVHOT_REQUIRE(d && out, AVHOT_ERR_INVALID, "tfm_create: null argument");
  const int C = d->input_dim, H = d->hidden_dim, O = d->output_dim, L = d->num_layers, F = d->ffn_dim;
  AVHOT_REQUIRE(C > 0 && H > 0 && O > 0 && F > 0 && d->num_heads > 0 && d->max_len > 0, AVHOT_ERR_INVALID,
                "tfm_create: non-positive dimension");
  AVHOT_REQUIRE(L >= 1 && L <= AVHOT_TFM_MAX_LAYERS, AVHOT_ERR_UNSUPPORTED, "tfm_create: num_layers=%d outside 1..%d", L,
                AVHOT_TFM_MAX_LAYERS);
  AVHOT_REQUIRE((C & 3) == 0 && (H & 7) == 0 && (F & 7) == 0 && H <= 1024, AVHOT_ERR_UNSUPPORTED,
                "tfm_create: input_dim %% 4, hidden_dim %% 8 (<= 1024) and ffn_dim %% 8 must be 0 (got %d, %d, %d)", C, H, F);
  AVHOT_REQUIRE(H % d->num_heads == 0 && H / d->num_heads <= SA_MAX_D, AVHOT_ERR_UNSUPPORTED,
                "tfm_create: head dim %d not supported", H / d->num_heads);
  AVHOT_REQUIRE(d->input_w && d->input_b && d->pos && d->proj_w && d->proj_b, AVHOT_ERR_INVALID, "tfm_create: tensors missing");
  for (int l = 0; l < L; ++l)
    AVHOT_REQUIRE(d->in_proj_w[l] && d->in_proj_b[l] && d->out_w[l] && d->out_b[l] && d->lin1_w[l] && d->lin1_b[l] &&
                      d->lin2_w[l] && d->lin2_b[l] && d->n1_w[l] && d->n1_b[l] && d->n2_w[l] && d->n2_b[l],
                  AVHOT_ERR_INVALID, "tfm_create: layer %d tensors missing", l);
  avhot_tfm* h = new avhot_tfm();
  h->in_dim = C; h->H = H; h->out_dim = O; h->layers = L; h->heads = d->num_heads; h->F = F; h->max_len = d->max_len;
  h->eps = d->ln_eps;
  struct Item { float** dst; const float* src; size_t n; };
  std::vector<Item> items = {{&h->input_w, d->input_w, (size_t)H * C}, {&h->input_b, d->input_b, (size_t)H},
                             {&h->pos, d->pos, (size_t)d->max_len * H}, {&h->proj_w, d->proj_w, (size_t)O * H},
                             {&h->proj_b, d->proj_b, (size_t)O}};
  for (int l = 0; l < L; ++l) {
    items.push_back({&h->in_proj_w[l], d->in_proj_w[l], (size_t)3 * H * H});
    items.push_back({&h->in_proj_b[l], d->in_proj_b[l], (size_t)3 * H});
    items.push_back({&h->out_w[l], d->out_w[l], (size_t)H * H});
    items.push_back({&h->out_b[l], d->out_b[l], (size_t)H});
    items.push_back({&h->lin1_w[l], d->lin1_w[l], (size_t)F * H});
    items.push_back({&h->lin1_b[l], d->lin1_b[l], (size_t)F});
    items.push_back({&h->lin2_w[l], d->lin2_w[l], (size_t)H * F});
    items.push_back({&h->lin2_b[l], d->lin2_b[l], (size_t)H});
    items.push_back({&h->n1_w[l], d->n1_w[l], (size_t)H});
    items.push_back({&h->n1_b[l], d->n1_b[l], (size_t)H});
    items.push_back({&h->n2_w[l], d->n2_w[l], (size_t)H});
    items.push_back({&h->n2_b[l], d->n2_b[l], (size_t)H});
  }
  size_t total = 0;
  for (const Item& it : items) total += salign(sizeof(float) * it.n);
  h->arena = nullptr;
  if (cudaMalloc(&h->arena, total) != cudaSuccess) {
    delete h;
    set_error("tfm_create: cudaMalloc(%zu) failed", total);
    return AVHOT_ERR_CUDA;
  }
  char* p = reinterpret_cast<char*>(h->arena);
  cudaStream_t s = as_stream(stream);
  for (const Item& it : items) {
    *it.dst = reinterpret_cast<float*>(p);
    if (cudaMemcpyAsync(p, it.src, sizeof(float) * it.n, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
      set_error("tfm_create: device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      cudaFree(h->arena);
      delete h;
      return AVHOT_ERR_CUDA;
    }
    p += salign(sizeof(float) * it.n);
  }
  *out = h;
  return AVHOT_OK;
}

extern "C" void avhot_tfm_destroy(avhot_tfm_t h) {
  if (!h) return;
  cudaFree(h->arena);
  delete h;
}

// scratch: x0, sa0 (pass-invariant), x, y, ctx (rows x H each), qkv (rows x 3H), ff (rows x F), pooled (B P x H)
extern "C" size_t avhot_tfm_workspace_bytes(avhot_tfm_t h, int B, int T, int passes) {
  if (!h || B <= 0 || T <= 0 || passes <= 0) return 0;
  const size_t rows = (size_t)B * T, H = h->H, F = h->F;
  return 5 * salign(sizeof(float) * rows * H) + salign(sizeof(float) * rows * 3 * H) + salign(sizeof(float) * rows * F) +
         salign(sizeof(float) * (size_t)B * passes * H);
}

extern "C" int avhot_tfm_forward(avhot_tfm_t h, const float* seq, const int32_t* lengths, float* z, int B, int T,
                                 int passes, float p, uint64_t seed, uint64_t clip0, uint32_t t0, int precision,
                                 void* workspace, size_t workspace_bytes, avhot_stream_t stream) {
  AVHOT_REQUIRE(h && seq && z && B >= 0 && T >= 1 && passes >= 1, AVHOT_ERR_INVALID, "tfm_forward: bad arguments");
  AVHOT_REQUIRE(precision == AVHOT_PREC_FP32 || precision == AVHOT_PREC_TC, AVHOT_ERR_UNSUPPORTED,
                "tfm_forward: precision %d not built", precision);
  AVHOT_REQUIRE(p >= 0.f && p < 1.f, AVHOT_ERR_INVALID, "tfm_forward: dropout p=%g outside [0, 1)", (double)p);
  if (B == 0) return AVHOT_OK;
  AVHOT_REQUIRE((long long)B * T <= 0x7fffffff / 8 && (long long)T * h->F <= 0x7fffffffLL, AVHOT_ERR_INVALID,
                "tfm_forward: B*T too large for one call");
  const size_t need = avhot_tfm_workspace_bytes(h, B, T, passes);
  AVHOT_REQUIRE(workspace && workspace_bytes >= need, AVHOT_ERR_WORKSPACE, "tfm_forward: workspace %zu < %zu bytes",
                workspace_bytes, need);
  const int H = h->H, F = h->F, d = H / h->heads;
  const long long rows = (long long)B * T;
  cudaStream_t s = as_stream(stream);
  char* q = reinterpret_cast<char*>(workspace);
  auto take = [&](size_t n) { float* r = reinterpret_cast<float*>(q); q += salign(sizeof(float) * n); return r; };
  float* x0 = take((size_t)rows * H);
  float* sa0 = take((size_t)rows * H);
  float* x = take((size_t)rows * H);
  float* y = take((size_t)rows * H);
  float* ctx = take((size_t)rows * H);
  float* qkv = take((size_t)rows * 3 * H);
  float* ff = take((size_t)rows * F);
  float* pooled = take((size_t)B * passes * H);
  const int M = (int)rows;
  const int ew_grid = num_sms() * 8;
  // every nn.Linear of the encoder: fp32 FMA (parity path) or the tcgen05 TF32 GEMM the video Linear uses (AVHOT_PREC_TC:
  // fp32 operands read as TF32, fp32 accumulate; attention, LayerNorm, GELU and the pooled projection stay fp32)
  const bool tc = precision == AVHOT_PREC_TC;
  auto gemm = [&](const float* a, int lda, const float* w, int ldw, const float* bias, float* c, int ldc, int m, int n, int k,
                  const char* name) -> int {
    return tc ? gemm_tf32_tc(a, lda, w, ldw, bias, c, ldc, m, n, k, ACT_NONE, s, name)
              : gemm_f32(a, nullptr, lda, w, ldw, bias, c, ldc, m, n, k, ACT_NONE, s, name);
  };

  // x0 = input_proj(seq) + pos   (src/encoders.py:172, :180-182)
  int rc = gemm(seq, h->in_dim, h->input_w, h->in_dim, h->input_b, x0, H, M, H, h->in_dim, "tfm_input_proj");
  if (rc != AVHOT_OK) return rc;
  {
    ProfScope prof("tfm_add_pos", s);
    add_pos_kernel<<<ew_grid, 256, 0, s>>>(x0, h->pos, rows, T, H, h->max_len);
    AVHOT_LAUNCH_CHECK();
  }
  // self-attention block of one layer: y_out = out_proj(MHA(x_in))
  auto sa_block = [&](int l, const float* x_in, float* y_out) -> int {
    int r = gemm(x_in, H, h->in_proj_w[l], H, h->in_proj_b[l], qkv, 3 * H, M, 3 * H, H, "tfm_qkv");
    if (r != AVHOT_OK) return r;
    r = launch_self_attn(qkv, 3 * H, 0, H, 2 * H, lengths, nullptr, nullptr, ctx, H, nullptr, B, T, h->heads, d, s,
                         "tfm_self_attn");
    if (r != AVHOT_OK) return r;
    return gemm(ctx, H, h->out_w[l], H, h->out_b[l], y_out, H, M, H, H, "tfm_out_proj");
  };
  rc = sa_block(0, x0, sa0);     // pass-invariant: every dropout of layer 0 comes after it
  if (rc != AVHOT_OK) return rc;
  const DropSpec ds = make_drop(p, seed, clip0, t0);
  const unsigned ln_grid = (unsigned)((rows + 7) / 8);
  for (int pass = 0; pass < passes; ++pass) {
    const float* x_in = x0;
    for (int l = 0; l < h->layers; ++l) {
      const float* sa = sa0;
      if (l > 0) {
        rc = sa_block(l, x_in, y);
        if (rc != AVHOT_OK) return rc;
        sa = y;
      }
      {
        ProfScope prof("tfm_resid_ln", s);
        resid_drop_ln_kernel<<<ln_grid, 256, 0, s>>>(x_in, sa, h->n1_w[l], h->n1_b[l], h->eps, x, rows, T, H,
                                                     AVHOT_SITE_TFM + 3u * (uint32_t)l, pass, ds);
        AVHOT_LAUNCH_CHECK();
      }
      rc = gemm(x, H, h->lin1_w[l], H, h->lin1_b[l], ff, F, M, F, H, "tfm_linear1");
      if (rc != AVHOT_OK) return rc;
      {
        ProfScope prof("tfm_gelu", s);
        gelu_drop_kernel<<<ew_grid, 256, 0, s>>>(ff, rows, T, F, AVHOT_SITE_TFM + 3u * (uint32_t)l + 1u, pass, ds);
        AVHOT_LAUNCH_CHECK();
      }
      rc = gemm(ff, F, h->lin2_w[l], F, h->lin2_b[l], y, H, M, H, F, "tfm_linear2");
      if (rc != AVHOT_OK) return rc;
      {
        ProfScope prof("tfm_resid_ln", s);
        resid_drop_ln_kernel<<<ln_grid, 256, 0, s>>>(x, y, h->n2_w[l], h->n2_b[l], h->eps, x, rows, T, H,
                                                     AVHOT_SITE_TFM + 3u * (uint32_t)l + 2u, pass, ds);
        AVHOT_LAUNCH_CHECK();
      }
      x_in = x;
    }
    ProfScope prof("tfm_pool", s);
    seq_mean_pool_kernel<<<B, 256, 0, s>>>(x, lengths, pooled, T, H, passes, pass);
    AVHOT_LAUNCH_CHECK();
  }
  return gemm_f32(pooled, nullptr, H, h->proj_w, H, h->proj_b, z, h->out_dim, B * passes, h->out_dim, H, ACT_NONE, s,
                  "tfm_projection");
}

// =================================================================================================================
// TemporalAttention
// =================================================================================================================
struct avhot_tattn {
  int feat, H, heads;
  float* arena;
  float *qkv_w, *qkv_b, *o_w, *o_b;   // q / k / v projections stacked (3H, feat): one GEMM
};

extern "C" int avhot_tattn_create(const avhot_tattn_desc* d, avhot_tattn_t* out, avhot_stream_t stream) {
  AVHOT_REQUIRE(d && out, AVHOT_ERR_INVALID, "tattn_create: null argument");
  const int C = d->feature_dim, H = d->hidden_dim;
  AVHOT_REQUIRE(C > 0 && H > 0 && d->num_heads > 0, AVHOT_ERR_INVALID, "tattn_create: non-positive dimension");
  AVHOT_REQUIRE((C & 3) == 0 && (H & 3) == 0, AVHOT_ERR_UNSUPPORTED,
                "tattn_create: feature_dim=%d and hidden_dim=%d must be multiples of 4", C, H);
  AVHOT_REQUIRE(H / d->num_heads >= 1 && H / d->num_heads <= SA_MAX_D, AVHOT_ERR_UNSUPPORTED,
                "tattn_create: head dim %d not supported", H / d->num_heads);
  AVHOT_REQUIRE(d->q_w && d->q_b && d->k_w && d->k_b && d->v_w && d->v_b && d->o_w && d->o_b, AVHOT_ERR_INVALID,
                "tattn_create: tensors missing");
  avhot_tattn* h = new avhot_tattn();
  h->feat = C; h->H = H; h->heads = d->num_heads;
  const size_t nw = (size_t)H * C;
  const size_t total = salign(sizeof(float) * 3 * nw) + salign(sizeof(float) * 3 * H) + salign(sizeof(float) * (size_t)H * H) +
                       salign(sizeof(float) * H);
  h->arena = nullptr;
  if (cudaMalloc(&h->arena, total) != cudaSuccess) {
    delete h;
    set_error("tattn_create: cudaMalloc(%zu) failed", total);
    return AVHOT_ERR_CUDA;
  }
  char* p = reinterpret_cast<char*>(h->arena);
  h->qkv_w = reinterpret_cast<float*>(p); p += salign(sizeof(float) * 3 * nw);
  h->qkv_b = reinterpret_cast<float*>(p); p += salign(sizeof(float) * 3 * H);
  h->o_w = reinterpret_cast<float*>(p); p += salign(sizeof(float) * (size_t)H * H);
  h->o_b = reinterpret_cast<float*>(p);
  cudaStream_t s = as_stream(stream);
  const float* ws[3] = {d->q_w, d->k_w, d->v_w};
  const float* bs[3] = {d->q_b, d->k_b, d->v_b};
  bool ok = true;
  for (int i = 0; i < 3; ++i) {
    ok = ok && cudaMemcpyAsync(h->qkv_w + i * nw, ws[i], sizeof(float) * nw, cudaMemcpyDeviceToDevice, s) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(h->qkv_b + (size_t)i * H, bs[i], sizeof(float) * H, cudaMemcpyDeviceToDevice, s) == cudaSuccess;
  }
  ok = ok && cudaMemcpyAsync(h->o_w, d->o_w, sizeof(float) * (size_t)H * H, cudaMemcpyDeviceToDevice, s) == cudaSuccess;
  ok = ok && cudaMemcpyAsync(h->o_b, d->o_b, sizeof(float) * H, cudaMemcpyDeviceToDevice, s) == cudaSuccess;
  if (!ok) {
    set_error("tattn_create: device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(h->arena);
    delete h;
    return AVHOT_ERR_CUDA;
  }
  *out = h;
  return AVHOT_OK;
}

extern "C" void avhot_tattn_destroy(avhot_tattn_t h) {
  if (!h) return;
  cudaFree(h->arena);
  delete h;
}

extern "C" size_t avhot_tattn_workspace_bytes(avhot_tattn_t h, int B, int S) {
  if (!h || B <= 0 || S <= 0) return 0;
  const size_t rows = (size_t)B * S;
  return salign(sizeof(float) * rows * 3 * h->H) + salign(sizeof(float) * rows * h->H);
}

extern "C" int avhot_tattn_forward(avhot_tattn_t h, const float* seq, const uint8_t* valid, float* attended,
                                   float* weights, int B, int S, void* workspace, size_t workspace_bytes,
                                   avhot_stream_t stream) {
  AVHOT_REQUIRE(h && seq && attended && B >= 0 && S >= 1, AVHOT_ERR_INVALID, "tattn_forward: bad arguments");
  if (B == 0) return AVHOT_OK;
  AVHOT_REQUIRE((long long)B * S <= 0x7fffffff / 8, AVHOT_ERR_INVALID, "tattn_forward: B*S too large for one call");
  const size_t need = avhot_tattn_workspace_bytes(h, B, S);
  AVHOT_REQUIRE(workspace && workspace_bytes >= need, AVHOT_ERR_WORKSPACE, "tattn_forward: workspace %zu < %zu bytes",
                workspace_bytes, need);
  const int H = h->H, M = B * S;
  cudaStream_t s = as_stream(stream);
  float* qkv = reinterpret_cast<float*>(workspace);
  float* ctx = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + salign(sizeof(float) * (size_t)M * 3 * H));
  int rc = gemm_f32(seq, nullptr, h->feat, h->qkv_w, h->feat, h->qkv_b, qkv, 3 * H, M, 3 * H, h->feat, ACT_NONE, s, "tattn_qkv");
  if (rc != AVHOT_OK) return rc;
  rc = launch_self_attn(qkv, 3 * H, 0, H, 2 * H, nullptr, valid, valid, ctx, H, weights, B, S, h->heads, H / h->heads, s,
                        "tattn_self_attn");
  if (rc != AVHOT_OK) return rc;
  return gemm_f32(ctx, nullptr, H, h->o_w, H, h->o_b, attended, H, M, H, H, ACT_NONE, s, "tattn_out_proj");
}

extern "C" int avhot_tattn_pool(const float* seq, const float* weights, float* pooled, int B, int S, int D, int heads,
                                avhot_stream_t stream) {
  AVHOT_REQUIRE(seq && weights && pooled && B >= 0 && S >= 1 && D >= 1 && heads >= 1, AVHOT_ERR_INVALID,
                "tattn_pool: bad arguments");
  AVHOT_REQUIRE(S <= 12000, AVHOT_ERR_UNSUPPORTED, "tattn_pool: S=%d > 12000", S);
  if (B == 0) return AVHOT_OK;
  cudaStream_t s = as_stream(stream);
  ProfScope prof("tattn_pool", s);
  tattn_pool_kernel<<<B, 256, sizeof(float) * S, s>>>(seq, weights, pooled, S, D, heads);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

// =================================================================================================================
// PairwiseModalityAttention aggregation
// =================================================================================================================
extern "C" int avhot_pair_combine(const float* self_feat, const float* const* msgs_host_array, int n_msgs,
                                  const uint8_t* avail, const float* ln_w, const float* ln_b, float ln_eps, float* out,
                                  int B, int H, float p, uint64_t seed, uint64_t clip0, uint32_t t, uint32_t site,
                                  avhot_stream_t stream) {
  AVHOT_REQUIRE(self_feat && ln_w && ln_b && out && B >= 0 && H >= 1, AVHOT_ERR_INVALID, "pair_combine: bad arguments");
  AVHOT_REQUIRE(n_msgs >= 0 && n_msgs <= 3 && (n_msgs == 0 || msgs_host_array), AVHOT_ERR_INVALID,
                "pair_combine: n_msgs=%d outside 0..3", n_msgs);
  AVHOT_REQUIRE((H & 7) == 0 && H <= 1024, AVHOT_ERR_UNSUPPORTED, "pair_combine: hidden_dim=%d must be a multiple of 8, <= 1024", H);
  AVHOT_REQUIRE(p >= 0.f && p < 1.f, AVHOT_ERR_INVALID, "pair_combine: dropout p=%g outside [0, 1)", (double)p);
  if (B == 0) return AVHOT_OK;
  const float* m[3] = {nullptr, nullptr, nullptr};
  for (int i = 0; i < n_msgs; ++i) {
    AVHOT_REQUIRE(msgs_host_array[i], AVHOT_ERR_INVALID, "pair_combine: message %d is NULL", i);
    m[i] = msgs_host_array[i];
  }
  cudaStream_t s = as_stream(stream);
  ProfScope prof("pair_combine", s);
  pair_combine_kernel<<<ceil_div(B, 8), 256, 0, s>>>(self_feat, m[0], m[1], m[2], avail, ln_w, ln_b, ln_eps, out, B, H, site,
                                                     make_drop(p, seed, clip0, t));
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}
