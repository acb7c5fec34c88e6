// SequenceEncoder(encoder_type='cnn')  (src/encoders.py:79-94 ctor, :157-168 forward) - the
// variant the reference's author ran locally (outputs/2025-12-01/20-55-09/.hydra/config.yaml:26).
//
//   seq (B, T, in) -> Conv1d(in -> H, k 5, pad 2) -> BatchNorm1d (eval: running statistics) -> ReLU -> Dropout
//       -> Conv1d(H -> H, k 3, pad 1) -> BatchNorm1d -> ReLU -> mean over T -> Dropout -> Linear(H -> out)
//
// fp32 path.  Each convolution is the fp32 GEMM of gemm.cu over a zero-padded copy of its input: with the input
// stored (B, T + 2 pad, C) row-major, the k taps of output position t are the k * C contiguous floats starting at
// padded row t, so A is that buffer read with lda = C < K = k * C (overlapping rows), W is the kernel reordered to
// (H, k * C) tap-major with the BatchNorm affine folded in, and ReLU is the GEMM's activation.  The few rows that
// straddle two clips are computed and ignored.  The same nn.Dropout module is applied twice (:161, :167): site
// AVHOT_SITE_CNN0 masks the (H, T) activation (element h * T + t, a dense site: 16-bit draws), AVHOT_SITE_CNN1 the
// pooled (H,) vector.  Everything after the first convolution depends on the pass, so P passes cost P second
// convolutions - this is the reference's arithmetic, not a fused hot path.
#include "common.cuh"

struct avhot_cnn {
  int in_dim, hidden_dim, output_dim;
  float *w1, *b1;   // (H, 5 * in) tap-major, BN1 folded; (H)
  float *w2, *b2;   // (H, 3 * H) tap-major, BN2 folded; (H)
  float *proj_w, *proj_b;
  float* arena;
};

namespace avhot {

constexpr int CK1 = 5, CK2 = 3;

// out[co][k * C + ci] = w[co][ci][k] * g[co],  bias_out[co] = (b[co] - mean[co]) * g[co] + beta[co],
// g = gamma / sqrt(var + eps)      (BatchNorm1d in eval mode folded into the convolution)
__global__ void fold_conv_bn_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                    float* __restrict__ w_out, float* __restrict__ b_out, int C, int KS) {
  const int co = blockIdx.x;
  const float g = gamma[co] / sqrtf(var[co] + eps);
  for (int i = threadIdx.x; i < C * KS; i += blockDim.x) {
    const int k = i / C, ci = i % C;
    w_out[(size_t)co * C * KS + i] = w[((size_t)co * C + ci) * KS + k] * g;
  }
  if (threadIdx.x == 0) b_out[co] = (b[co] - mean[co]) * g + beta[co];
}

// x (B, T, C) -> xp (B, T + 2 pad, C) with zero rows at both ends (+ `tail` zero rows after the last clip)
__global__ void pad_rows_kernel(const float* __restrict__ x, float* __restrict__ xp, int B, int T, int C, int pad, int tail) {
  const long long n4 = ((long long)B * (T + 2 * pad) + tail) * (C / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / (C / 4);
    const int c4 = (int)(i % (C / 4));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < (long long)B * (T + 2 * pad)) {
      const int b = (int)(row / (T + 2 * pad)), t = (int)(row % (T + 2 * pad)) - pad;
      if (t >= 0 && t < T) v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)b * T + t) * C) + c4);
    }
    reinterpret_cast<float4*>(xp)[i] = v;
  }
}

// a1: conv1 output rows (b, t) at row b * (T + 4) + t, (.., H) -> yp (B, T + 2, H): dropout (site CNN0, element
// h * T + t of clip b) then zero padding.  One thread per Philox block of 8 elements.
__global__ void cnn_drop_pad_kernel(const float* __restrict__ a1, float* __restrict__ yp, int B, int T, int H,
                                    int pass, DropSpec ds) {
  const long long per_clip = ((long long)H * T + 7) / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * per_clip) return;
  const int b = (int)(i / per_clip);
  const uint32_t blk = (uint32_t)(i % per_clip);
  const bool drop = ds.thr != 0u;
  Philox4 r = {0, 0, 0, 0};
  if (drop) r = philox4x32_10(blk, (uint32_t)(ds.clip0 + (uint64_t)b), ds.t0 + (uint32_t)pass, AVHOT_SITE_CNN0, ds.k0, ds.k1);
  const uint32_t thr16 = ds.thr >> 16;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const long long e = (long long)blk * 8 + q;
    if (e >= (long long)H * T) break;
    const int h = (int)(e / T), t = (int)(e % T);
    float v = a1[((size_t)b * (T + 2 * 2) + t) * H + h];
    if (drop) v = philox_half(r, (uint32_t)q) >= thr16 ? v * ds.scale : 0.f;
    yp[((size_t)b * (T + 2) + t + 1) * H + h] = v;
  }
}
__global__ void cnn_zero_pad_rows_kernel(float* __restrict__ yp, int B, int T, int H, int tail) {
  // rows 0 and T + 1 of every clip, and `tail` rows after the last clip
  const long long n = ((long long)2 * B + tail) * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / H;
    const int h = (int)(i % H);
    size_t row;
    if (r < 2LL * B) row = (size_t)(r / 2) * (T + 2) + ((r & 1) ? T + 1 : 0);
    else row = (size_t)B * (T + 2) + (size_t)(r - 2LL * B);
    yp[row * H + h] = 0.f;
  }
}

// a2: conv2 output rows (b, t) at row b * (T + 2) + t -> pooled[(b * P + pass)][h] = dropout(mean_t a2), site CNN1
__global__ void cnn_pool_drop_kernel(const float* __restrict__ a2, float* __restrict__ pooled, int B, int T, int H,
                                     int P, int pass, DropSpec ds) {
  const int b = blockIdx.x;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += a2[((size_t)b * (T + 2) + t) * H + h];
    float v = s / (float)T;
    if (ds.thr != 0u) {
      const Philox4 w = philox4x32_10((uint32_t)h >> 2, (uint32_t)(ds.clip0 + (uint64_t)b), ds.t0 + (uint32_t)pass,
                                      AVHOT_SITE_CNN1, ds.k0, ds.k1);
      v = philox_word(w, h & 3) >= ds.thr ? v * ds.scale : 0.f;
    }
    pooled[((size_t)b * P + pass) * H + h] = v;
  }
}

static size_t calign(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace avhot

using namespace avhot;

extern "C" int avhot_cnn_create(const avhot_cnn_desc* d, avhot_cnn_t* out, avhot_stream_t stream) {
  AVHOT_REQUIRE(d && out, AVHOT_ERR_INVALID, "cnn_create: null argument");
  AVHOT_REQUIRE(d->input_dim > 0 && d->hidden_dim > 0 && d->output_dim > 0 && (d->input_dim & 3) == 0 &&
                    (d->hidden_dim & 3) == 0, AVHOT_ERR_UNSUPPORTED,
                "cnn_create: input_dim=%d and hidden_dim=%d must be positive multiples of 4", d->input_dim, d->hidden_dim);
  AVHOT_REQUIRE(d->conv1_weight && d->conv1_bias && d->bn1_weight && d->bn1_bias && d->bn1_mean && d->bn1_var &&
                    d->conv2_weight && d->conv2_bias && d->bn2_weight && d->bn2_bias && d->bn2_mean && d->bn2_var &&
                    d->proj_weight && d->proj_bias, AVHOT_ERR_INVALID, "cnn_create: tensors missing");
  const int C = d->input_dim, H = d->hidden_dim, O = d->output_dim;
  const size_t total = calign(sizeof(float) * (size_t)H * CK1 * C) + calign(sizeof(float) * (size_t)H * CK2 * H) +
                       2 * calign(sizeof(float) * H) + calign(sizeof(float) * (size_t)O * H) + calign(sizeof(float) * O);
  avhot_cnn* h = new avhot_cnn();
  h->in_dim = C; h->hidden_dim = H; h->output_dim = O; h->arena = nullptr;
  if (cudaMalloc(&h->arena, total) != cudaSuccess) {
    delete h;
    set_error("cnn_create: cudaMalloc(%zu) failed", total);
    return AVHOT_ERR_CUDA;
  }
  cudaStream_t s = as_stream(stream);
  char* p = reinterpret_cast<char*>(h->arena);
  auto take = [&](size_t bytes) { float* r = reinterpret_cast<float*>(p); p += calign(bytes); return r; };
  h->w1 = take(sizeof(float) * (size_t)H * CK1 * C); h->b1 = take(sizeof(float) * H);
  h->w2 = take(sizeof(float) * (size_t)H * CK2 * H); h->b2 = take(sizeof(float) * H);
  h->proj_w = take(sizeof(float) * (size_t)O * H); h->proj_b = take(sizeof(float) * O);
  fold_conv_bn_kernel<<<H, 128, 0, s>>>(d->conv1_weight, d->conv1_bias, d->bn1_weight, d->bn1_bias, d->bn1_mean,
                                        d->bn1_var, d->bn_eps, h->w1, h->b1, C, CK1);
  fold_conv_bn_kernel<<<H, 128, 0, s>>>(d->conv2_weight, d->conv2_bias, d->bn2_weight, d->bn2_bias, d->bn2_mean,
                                        d->bn2_var, d->bn_eps, h->w2, h->b2, H, CK2);
  const bool ok = cudaGetLastError() == cudaSuccess &&
                  cudaMemcpyAsync(h->proj_w, d->proj_weight, sizeof(float) * (size_t)O * H, cudaMemcpyDeviceToDevice, s) == cudaSuccess &&
                  cudaMemcpyAsync(h->proj_b, d->proj_bias, sizeof(float) * O, cudaMemcpyDeviceToDevice, s) == cudaSuccess;
  if (!ok) {
    set_error("cnn_create: device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(h->arena);
    delete h;
    return AVHOT_ERR_CUDA;
  }
  *out = h;
  return AVHOT_OK;
}

extern "C" void avhot_cnn_destroy(avhot_cnn_t h) {
  if (!h) return;
  cudaFree(h->arena);
  delete h;
}

// scratch: padded input, conv1 output, padded dropped activation, conv2 output, pooled rows
extern "C" size_t avhot_cnn_workspace_bytes(avhot_cnn_t h, int B, int T, int passes) {
  if (!h || B <= 0 || T <= 0 || passes <= 0) return 0;
  const size_t C = h->in_dim, H = h->hidden_dim;
  return calign(sizeof(float) * ((size_t)B * (T + 4) + CK1) * C) + calign(sizeof(float) * (size_t)B * (T + 4) * H) +
         calign(sizeof(float) * ((size_t)B * (T + 2) + CK2) * H) + calign(sizeof(float) * (size_t)B * (T + 2) * H) +
         calign(sizeof(float) * (size_t)B * passes * H);
}

extern "C" int avhot_cnn_forward(avhot_cnn_t h, const float* seq, float* z, int B, int T, int passes, float p,
                                 uint64_t seed, uint64_t clip0, uint32_t t0, void* workspace, size_t workspace_bytes,
                                 avhot_stream_t stream) {
  AVHOT_REQUIRE(h && seq && z && B >= 0 && T >= 1 && passes >= 1, AVHOT_ERR_INVALID, "cnn_forward: bad arguments");
  AVHOT_REQUIRE(p >= 0.f && p < 1.f, AVHOT_ERR_INVALID, "cnn_forward: dropout p=%g outside [0, 1)", (double)p);
  if (B == 0) return AVHOT_OK;
  AVHOT_REQUIRE((long long)B * (T + 4) <= 0x7fffffff / 8, AVHOT_ERR_INVALID, "cnn_forward: B*T too large for one call");
  const size_t need = avhot_cnn_workspace_bytes(h, B, T, passes);
  AVHOT_REQUIRE(workspace && workspace_bytes >= need, AVHOT_ERR_WORKSPACE, "cnn_forward: workspace %zu < %zu bytes",
                workspace_bytes, need);
  const int C = h->in_dim, H = h->hidden_dim;
  cudaStream_t s = as_stream(stream);
  char* q = reinterpret_cast<char*>(workspace);
  float* xp = reinterpret_cast<float*>(q); q += calign(sizeof(float) * ((size_t)B * (T + 4) + CK1) * C);
  float* a1 = reinterpret_cast<float*>(q); q += calign(sizeof(float) * (size_t)B * (T + 4) * H);
  float* yp = reinterpret_cast<float*>(q); q += calign(sizeof(float) * ((size_t)B * (T + 2) + CK2) * H);
  float* a2 = reinterpret_cast<float*>(q); q += calign(sizeof(float) * (size_t)B * (T + 2) * H);
  float* pooled = reinterpret_cast<float*>(q);
  {
    ProfScope prof("cnn_pad", s);
    pad_rows_kernel<<<num_sms() * 4, 256, 0, s>>>(seq, xp, B, T, C, 2, CK1);
    AVHOT_LAUNCH_CHECK();
  }
  int rc = gemm_f32(xp, nullptr, C, h->w1, CK1 * C, h->b1, a1, H, B * (T + 4), H, CK1 * C, ACT_RELU, s, "cnn_conv1");
  if (rc != AVHOT_OK) return rc;
  {
    ProfScope prof("cnn_pad", s);
    cnn_zero_pad_rows_kernel<<<ceil_div((2 * B + CK2) * H, 256), 256, 0, s>>>(yp, B, T, H, CK2);
    AVHOT_LAUNCH_CHECK();
  }
  const DropSpec ds = make_drop(p, seed, clip0, t0);
  const long long blocks = (long long)B * (((long long)H * T + 7) / 8);
  for (int pass = 0; pass < passes; ++pass) {
    {
      ProfScope prof("cnn_dropout", s);
      cnn_drop_pad_kernel<<<(unsigned)((blocks + 255) / 256), 256, 0, s>>>(a1, yp, B, T, H, pass, ds);
      AVHOT_LAUNCH_CHECK();
    }
    rc = gemm_f32(yp, nullptr, H, h->w2, CK2 * H, h->b2, a2, H, B * (T + 2), H, CK2 * H, ACT_RELU, s, "cnn_conv2");
    if (rc != AVHOT_OK) return rc;
    ProfScope prof("cnn_pool", s);
    cnn_pool_drop_kernel<<<B, 256, 0, s>>>(a2, pooled, B, T, H, passes, pass, ds);
    AVHOT_LAUNCH_CHECK();
  }
  return gemm_f32(pooled, nullptr, H, h->proj_w, H, h->proj_b, z, h->output_dim, B * passes, h->output_dim, H, ACT_NONE,
                  s, "cnn_projection");
}
