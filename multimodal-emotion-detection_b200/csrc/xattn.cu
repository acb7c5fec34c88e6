// a5: CrossModalAttention for general shapes (src/attention.py:31-51 ctor, :80-141 forward).
// The hot path's two instances (Tq = 1, Tk = 2) are fused into fusion.cu; this is the standalone
// module: three input LayerNorms -> Q/K/V projections (gemm.cu) -> per (clip, head) scaled dot-product
// scores, key mask (-1e4), softmax, all-keys-masked rows to 0, Philox dropout on the weights, PV ->
// out_proj.  Also holds avhot_mc_reduce (a8, src/uncertainty.py:74-84).
#include "common.cuh"

struct avhot_xattn {
  int query_dim, key_dim, hidden_dim, num_heads;
  float *q_ln_w, *q_ln_b, *k_ln_w, *k_ln_b, *v_ln_w, *v_ln_b;
  float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b;
  float* arena;
};

namespace avhot {

// one warp per row: y = LayerNorm(x) * w + b (biased variance, eps 1e-5)
__global__ void layernorm_rows_kernel(const float* __restrict__ x, float* __restrict__ y,
                                      const float* __restrict__ w, const float* __restrict__ b,
                                      long long rows, int dim) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (size_t)row * dim;
  float s = 0.f;
  for (int c = lane; c < dim; c += 32) s += xr[c];
  const float mean = warp_sum(s) / (float)dim;
  float sq = 0.f;
  for (int c = lane; c < dim; c += 32) { const float d = xr[c] - mean; sq = fmaf(d, d, sq); }
  const float rstd = 1.f / sqrtf(warp_sum(sq) / (float)dim + 1e-5f);
  float* yr = y + (size_t)row * dim;
  for (int c = lane; c < dim; c += 32) yr[c] = (xr[c] - mean) * rstd * w[c] + b[c];
}

constexpr int XA_MAX_TK = 64;

// grid (heads, B); one warp per query row, lanes over keys (2 keys per lane) and then over head columns
__global__ void __launch_bounds__(128)
xattn_core_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                  const uint8_t* __restrict__ invalid, float* __restrict__ ctx, float* __restrict__ attn,
                  int Tq, int Tk, int D, int heads, uint32_t site, DropSpec ds) {
  extern __shared__ __align__(16) float xa_smem[];
  const int h = blockIdx.x, b = blockIdx.y, dh = D / heads;
  float* Ks = xa_smem;                 // [Tk][dh + 1]
  float* Vs = Ks + Tk * (dh + 1);      // [Tk][dh]
  for (int idx = threadIdx.x; idx < Tk * dh; idx += blockDim.x) {
    const int tk = idx / dh, d = idx % dh;
    Ks[tk * (dh + 1) + d] = K[((size_t)b * Tk + tk) * D + h * dh + d];
    Vs[tk * dh + d] = V[((size_t)b * Tk + tk) * D + h * dh + d];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float scale = 1.f / sqrtf((float)dh);
  bool all_inv = invalid != nullptr;
  if (invalid)
    for (int tk = 0; tk < Tk; ++tk) all_inv = all_inv && invalid[(size_t)b * Tk + tk];
  for (int tq = warp; tq < Tq; tq += 4) {
    const float* q = Q + ((size_t)b * Tq + tq) * D + h * dh;
    float s[2], a[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int tk = lane + 32 * u;
      float acc = -INFINITY;
      if (tk < Tk) {
        acc = 0.f;
        for (int d = 0; d < dh; ++d) acc = fmaf(__ldg(q + d), Ks[tk * (dh + 1) + d], acc);
        acc *= scale;                                                       // src/attention.py:112
        if (invalid && invalid[(size_t)b * Tk + tk]) acc = -1e4f;           // :115-118
      }
      s[u] = acc;
    }
    const float mx = warp_max(fmaxf(s[0], s[1]));
    const float e0 = lane < Tk ? expf(s[0] - mx) : 0.f, e1 = lane + 32 < Tk ? expf(s[1] - mx) : 0.f;
    const float den = warp_sum(e0 + e1);
    a[0] = e0 / den; a[1] = e1 / den;                                       // :121
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int tk = lane + 32 * u;
      if (tk >= Tk) { a[u] = 0.f; continue; }
      if (all_inv) a[u] = 0.f;                                              // :124-128
      if (ds.thr != 0u) {                                                   // :130
        const uint32_t e = (uint32_t)((h * Tq + tq) * Tk + tk);
        const Philox4 w = philox4x32_10(e >> 2, ds_clip(ds, b), ds.t0, site, ds.k0, ds.k1);
        a[u] = philox_word(w, e & 3) >= ds.thr ? a[u] * ds.scale : 0.f;
      }
      if (attn) attn[(((size_t)b * heads + h) * Tq + tq) * Tk + tk] = a[u];
    }
    for (int d = lane; d < dh; d += 32) {                                   // :133
      float acc = 0.f;
      for (int tk = 0; tk < Tk; ++tk) {
        const float w = __shfl_sync(0xffffffffu, tk < 32 ? a[0] : a[1], tk & 31);
        acc = fmaf(w, Vs[tk * dh + d], acc);
      }
      ctx[((size_t)b * Tq + tq) * D + h * dh + d] = acc;
    }
  }
}

// a8: src/uncertainty.py:74-84.  One warp per clip.
__global__ void mc_reduce_kernel(const float* __restrict__ logits, float* __restrict__ mean_logits,
                                 float* __restrict__ uncertainty, int32_t* __restrict__ labels,
                                 float* __restrict__ packed, int B, int T, int C) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  // lanes over classes (C <= 32); two passes over T: mean of probabilities, then variance
  const float* lg = logits + (size_t)b * T * C;
  float sum_logit = 0.f, sum_p = 0.f;
  for (int t = 0; t < T; ++t) {
    const float x = lane < C ? lg[(size_t)t * C + lane] : -INFINITY;
    const float mx = warp_max(x);
    const float e = lane < C ? expf(x - mx) : 0.f;
    const float p = e / warp_sum(e);
    sum_p += p;
    if (lane < C) sum_logit += x;
  }
  const float mean_p = sum_p / (float)T, mean_l = sum_logit / (float)T;     // :77
  float var = 0.f;
  for (int t = 0; t < T; ++t) {
    const float x = lane < C ? lg[(size_t)t * C + lane] : -INFINITY;
    const float mx = warp_max(x);
    const float e = lane < C ? expf(x - mx) : 0.f;
    const float p = e / warp_sum(e);
    const float d = p - mean_p;
    var = fmaf(d, d, var);
  }
  var = lane < C ? var / (float)T : 0.f;                                     // population variance, :81
  const float unc = warp_sum(var) / (float)C;                                // :84
  // argmax of the mean logits, first maximum wins (torch.argmax)
  float best = lane < C ? mean_l : -INFINITY;
  int idx = lane < C ? lane : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
  }
  if (lane < C) {
    if (mean_logits) mean_logits[(size_t)b * C + lane] = mean_l;
    if (packed) packed[(size_t)b * (C + 1) + lane] = mean_l;
  }
  if (lane == 0) {
    if (uncertainty) uncertainty[b] = unc;
    if (labels) labels[b] = idx;
    if (packed) packed[(size_t)b * (C + 1) + C] = unc;
  }
}

static size_t alx(size_t n_floats) { return (n_floats * sizeof(float) + 255) & ~(size_t)255; }

}  // namespace avhot

using namespace avhot;

extern "C" int avhot_xattn_create(const avhot_xattn_desc* d, avhot_xattn_t* out, avhot_stream_t stream) {
  AVHOT_REQUIRE(d && out, AVHOT_ERR_INVALID, "xattn_create: null argument");
  AVHOT_REQUIRE(d->num_heads >= 1 && d->hidden_dim > 0 && d->hidden_dim % d->num_heads == 0, AVHOT_ERR_INVALID,
                "xattn_create: hidden_dim %d must be divisible by num_heads %d (src/attention.py:34)",
                d->hidden_dim, d->num_heads);
  AVHOT_REQUIRE(d->query_dim > 0 && d->key_dim > 0 && !(d->query_dim & 3) && !(d->key_dim & 3) &&
                    !(d->hidden_dim & 3), AVHOT_ERR_UNSUPPORTED, "xattn_create: dims must be multiples of 4");
  AVHOT_REQUIRE(d->hidden_dim / d->num_heads <= 256, AVHOT_ERR_UNSUPPORTED, "xattn_create: head_dim > 256");
  const float* src[14] = {d->q_ln_w, d->q_ln_b, d->k_ln_w, d->k_ln_b, d->v_ln_w, d->v_ln_b, d->q_w, d->q_b,
                          d->k_w, d->k_b, d->v_w, d->v_b, d->o_w, d->o_b};
  const size_t D = d->hidden_dim, Dq = d->query_dim, Dk = d->key_dim;
  const size_t n[14] = {Dq, Dq, Dk, Dk, Dk, Dk, D * Dq, D, D * Dk, D, D * Dk, D, D * D, D};
  size_t total = 0;
  for (int i = 0; i < 14; ++i) {
    AVHOT_REQUIRE(src[i], AVHOT_ERR_INVALID, "xattn_create: tensor %d missing", i);
    total += alx(n[i]);
  }
  avhot_xattn* h = new avhot_xattn();
  h->query_dim = d->query_dim; h->key_dim = d->key_dim; h->hidden_dim = d->hidden_dim; h->num_heads = d->num_heads;
  cudaError_t e = cudaMalloc(&h->arena, total);
  if (e != cudaSuccess) {
    delete h;
    set_error("xattn_create: cudaMalloc(%zu) -> %s", total, cudaGetErrorString(e));
    return AVHOT_ERR_CUDA;
  }
  float** dst[14] = {&h->q_ln_w, &h->q_ln_b, &h->k_ln_w, &h->k_ln_b, &h->v_ln_w, &h->v_ln_b, &h->q_w, &h->q_b,
                     &h->k_w, &h->k_b, &h->v_w, &h->v_b, &h->o_w, &h->o_b};
  char* p = reinterpret_cast<char*>(h->arena);
  for (int i = 0; i < 14; ++i) {
    *dst[i] = reinterpret_cast<float*>(p);
    p += alx(n[i]);
    if (cudaMemcpyAsync(*dst[i], src[i], n[i] * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)) !=
        cudaSuccess) {
      set_error("xattn_create: device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      cudaFree(h->arena);
      delete h;
      return AVHOT_ERR_CUDA;
    }
  }
  *out = h;
  return AVHOT_OK;
}

extern "C" void avhot_xattn_destroy(avhot_xattn_t h) {
  if (!h) return;
  cudaFree(h->arena);
  delete h;
}

extern "C" size_t avhot_xattn_workspace_bytes(avhot_xattn_t h, int B, int Tq, int Tk) {
  if (!h || B <= 0 || Tq <= 0 || Tk <= 0) return 0;
  const size_t rq = (size_t)B * Tq, rk = (size_t)B * Tk, D = h->hidden_dim;
  return alx(rq * h->query_dim) + 2 * alx(rk * h->key_dim) + alx(rq * D) + 2 * alx(rk * D) + alx(rq * D);
}

extern "C" int avhot_xattn_forward(avhot_xattn_t h, const float* query, const float* key, const float* value,
                                   const uint8_t* invalid, float* out, float* attn, int B, int Tq, int Tk,
                                   float p, uint64_t seed, uint64_t clip0, uint32_t t, uint32_t site_index,
                                   void* workspace, size_t workspace_bytes, avhot_stream_t stream) {
  AVHOT_REQUIRE(h && query && key && value && out && B >= 0 && Tq >= 1 && Tk >= 1, AVHOT_ERR_INVALID,
                "xattn_forward: bad arguments");
  AVHOT_REQUIRE(Tk <= XA_MAX_TK, AVHOT_ERR_UNSUPPORTED, "xattn_forward: Tk=%d > %d", Tk, XA_MAX_TK);
  AVHOT_REQUIRE(p >= 0.f && p < 1.f, AVHOT_ERR_INVALID, "xattn_forward: dropout p=%g outside [0, 1)", (double)p);
  AVHOT_REQUIRE(B <= 65535, AVHOT_ERR_INVALID, "xattn_forward: B > 65535 per call");
  if (B == 0) return AVHOT_OK;
  const size_t need = avhot_xattn_workspace_bytes(h, B, Tq, Tk);
  AVHOT_REQUIRE(workspace && workspace_bytes >= need, AVHOT_ERR_WORKSPACE, "xattn_forward: workspace %zu < %zu bytes",
                workspace_bytes, need);
  cudaStream_t s = as_stream(stream);
  const size_t rq = (size_t)B * Tq, rk = (size_t)B * Tk;
  const int D = h->hidden_dim, Dq = h->query_dim, Dk = h->key_dim;
  char* w = reinterpret_cast<char*>(workspace);
  auto take = [&](size_t n_floats) { float* r = reinterpret_cast<float*>(w); w += alx(n_floats); return r; };
  float* qn = take(rq * Dq); float* kn = take(rk * Dk); float* vn = take(rk * Dk);
  float* Q = take(rq * D); float* K = take(rk * D); float* V = take(rk * D); float* ctx = take(rq * D);
  layernorm_rows_kernel<<<(unsigned)((rq + 3) / 4), 128, 0, s>>>(query, qn, h->q_ln_w, h->q_ln_b, (long long)rq, Dq);
  layernorm_rows_kernel<<<(unsigned)((rk + 3) / 4), 128, 0, s>>>(key, kn, h->k_ln_w, h->k_ln_b, (long long)rk, Dk);
  layernorm_rows_kernel<<<(unsigned)((rk + 3) / 4), 128, 0, s>>>(value, vn, h->v_ln_w, h->v_ln_b, (long long)rk, Dk);
  AVHOT_LAUNCH_CHECK();
  int rc = gemm_f32(qn, nullptr, Dq, h->q_w, Dq, h->q_b, Q, D, (int)rq, D, Dq, ACT_NONE, s);
  if (rc == AVHOT_OK) rc = gemm_f32(kn, nullptr, Dk, h->k_w, Dk, h->k_b, K, D, (int)rk, D, Dk, ACT_NONE, s);
  if (rc == AVHOT_OK) rc = gemm_f32(vn, nullptr, Dk, h->v_w, Dk, h->v_b, V, D, (int)rk, D, Dk, ACT_NONE, s);
  if (rc != AVHOT_OK) return rc;
  const int dh = D / h->num_heads;
  const size_t smem = sizeof(float) * (size_t)Tk * (2 * dh + 1);
  AVHOT_CUDA(cudaFuncSetAttribute(xattn_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  xattn_core_kernel<<<dim3(h->num_heads, B), 128, smem, s>>>(Q, K, V, invalid, ctx, attn, Tq, Tk, D, h->num_heads,
                                                               AVHOT_SITE_ATTN + site_index,
                                                               make_drop(p, seed, clip0, t));
  AVHOT_LAUNCH_CHECK();
  return gemm_f32(ctx, nullptr, D, h->o_w, D, h->o_b, out, D, (int)rq, D, D, ACT_NONE, s);
}

extern "C" int avhot_mc_reduce(const float* logits, float* mean_logits, float* uncertainty, int32_t* labels,
                               float* packed, int B, int T, int C, avhot_stream_t stream) {
  AVHOT_REQUIRE(logits && B >= 0 && T >= 1 && C >= 1 && C <= 32, AVHOT_ERR_INVALID,
                "mc_reduce: bad arguments (1 <= C <= 32)");
  if (B == 0) return AVHOT_OK;
  ProfScope prof("mc_reduce", as_stream(stream));
  mc_reduce_kernel<<<ceil_div(B, 4), 128, 0, as_stream(stream)>>>(logits, mean_logits, uncertainty, labels, packed,
                                                                 B, T, C);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}
