// a4: FrameEncoder(temporal_pooling='attention')  (src/encoders.py:203-251 ctor, :253-297 forward,
// :300-327 attention_pool).
//
//   frames (B, F, frame_dim) -> Linear + ReLU -> [Dropout] -> scores = Linear(H -> 1) -> softmax over F
//       -> weighted sum -> [Dropout] -> LayerNorm(H) -> Linear(H -> out)              = (B, out)
//
// Split at the only place the stochastic passes differ:
//   avhot_frame_hidden   Linear + ReLU, pass-invariant, the 102.8 MFLOP/clip contraction (gemm.cu);
//   avhot_frame_tail     everything after the ReLU for T passes.  One CTA = 32 rows (clip, pass);
//                        one warp per row streams the (F, H) activations once with an online
//                        softmax (Philox keep bits drawn in registers, lane <-> 8 consecutive columns =
//                        one Philox block of eight 16-bit draws per frame), then LayerNorm by warp shuffles;
//                        the 32 normalised rows are staged in shared memory and the H -> out
//                        projection runs as one 32-row tile so the weights are read once per CTA.
#include <float.h>

#include "common.cuh"

struct avhot_frame {
  int frame_dim, hidden_dim, output_dim;
  float* mlp_w;     // (H, frame_dim)
  float* mlp_b;     // (H)
  float* attn_w;    // (H)
  float* attn_b;    // (1)
  float* ln_w;      // (H)
  float* ln_b;      // (H)
  float* proj_wt;   // [j][o] = projection.1.weight[o][j]
  float* proj_b;    // (out)
  float* arena;
  bool has_attention;
  void* wq_f16;     // (H, frame_dim) fp16 = W s / 255 for the uint8 tensor-core GEMM (gemm_u8.cu), or NULL
  float* wq_inv;    // 1 / s (device)
};

namespace avhot {

constexpr int FH = 256;         // hidden size this kernel is specialised for
constexpr int FT_ROWS = 32;     // rows (clip, pass) per CTA
constexpr int FT_WARPS = 8;

// PROJECT = false stops after the LayerNorm and writes the normalised rows (B, T, H): the tensor-core hot path folds
// projection[1] into HybridFusion's proj_layers.video (two linear maps with nothing in between, src/encoders.py:249 ->
// src/fusion.py:486), so the H -> out projection - 31 % of this kernel's instructions - runs as part of a GEMM that is
// on the tensor cores anyway (avhot_frame_tail_norm).
// POOL: 0 = attention pooling (the hot path, configs/base.yaml:43), 1 = average, 2 = max (src/encoders.py:276-293).
// RAW (with PROJECT = false): write the pooled row before the second dropout / LayerNorm = FrameEncoder.attention_pool
// (src/encoders.py:300-327).
template <bool PROJECT, int POOL = 0, bool RAW = false>
__global__ void __launch_bounds__(FT_WARPS * 32, PROJECT ? 3 : 4)
frame_tail_kernel(const float* __restrict__ hid,        // (B, F, H) post-ReLU, pre-dropout
                  const uint8_t* __restrict__ valid,    // (B, F) 1 = valid frame, or NULL
                  float* __restrict__ z,                // (B, T, out)
                  const float* __restrict__ attn_w, const float* __restrict__ attn_b,
                  const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                  const float* __restrict__ proj_wt, const float* __restrict__ proj_b,
                  int B, int F, int T, int out_dim, DropSpec ds) {
  __shared__ __align__(16) float ys[FH][FT_ROWS];   // normalised rows, [j][r]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long n_rows = (long long)B * T;
  const long long row0 = (long long)blockIdx.x * FT_ROWS;
  const int c0 = 8 * lane, c1 = 8 * lane + 4;     // this lane's eight columns (two float4 groups)

  const float4 aw0 = __ldg(reinterpret_cast<const float4*>(attn_w + c0));
  const float4 aw1 = __ldg(reinterpret_cast<const float4*>(attn_w + c1));
  const float ab = __ldg(attn_b);
  const bool drop = ds.thr != 0u;
  const uint32_t thr16_hi = ds.thr & 0xFFFF0000u;      // (thr >> 16) << 16

  for (int i = warp; i < FT_ROWS; i += FT_WARPS) {
    const long long row = row0 + i;
    float y[8];
    if (row < n_rows) {                                    // warp-uniform
      // 32-bit row arithmetic (the entry points check B * T < 2^31): a 64-bit division costs ~100 instructions
      const int b = (int)((uint32_t)row / (uint32_t)T), t = (int)((uint32_t)row - (uint32_t)b * (uint32_t)T);
      const uint32_t g = ds_clip(ds, b), pass = ds.t0 + (uint32_t)t;
      const float* hb = hid + (size_t)b * F * FH;
      float m = -INFINITY, denom = 0.f;
      float pooled[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) pooled[q] = 0.f;
      for (int f = 0; f < F; ++f) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(hb + (size_t)f * FH + c0));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(hb + (size_t)f * FH + c1));
        float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        // site FRAME_MLP (dense: 16-bit draws): element e = f*H + col, block e >> 3 = f*32 + lane.  Dropped elements
        // become 0 here; the common factor 1 / (1 - p) of the kept ones is applied once to the score and once to the
        // pooled row instead of to each of the F * H elements.
        if (drop) {
          const Philox4 r = philox4x32_10_rk((uint32_t)(f * 32 + lane), g, pass, AVHOT_SITE_FRAME_MLP, ds);
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = philox_half_keep(r, (uint32_t)q, thr16_hi) ? x[q] : 0.f;
        }
        if (POOL != 0) {
          const bool ok = !valid || valid[(size_t)b * F + f];
          if (POOL == 1) {            // masked mean: sum of the valid frames (:277-283)
            if (ok) {
              denom += 1.f;
#pragma unroll
              for (int q = 0; q < 8; ++q) pooled[q] += x[q];
            }
          } else {                    // masked max: invalid frames count as finfo.min (:285-292)
#pragma unroll
            for (int q = 0; q < 8; ++q) pooled[q] = f == 0 ? (ok ? x[q] : -FLT_MAX) : fmaxf(pooled[q], ok ? x[q] : -FLT_MAX);
          }
          continue;
        }
        float s = x[0] * aw0.x + x[1] * aw0.y + x[2] * aw0.z + x[3] * aw0.w +
                  x[4] * aw1.x + x[5] * aw1.y + x[6] * aw1.z + x[7] * aw1.w;
        s = fmaf(warp_sum(s), ds.scale, ab);
        if (valid && !valid[(size_t)b * F + f]) s = -FLT_MAX;       // masked_fill(finfo.min), :322-324
        // online softmax; the running maximum changes ~log(F) times, so the rescale of the pooled row sits in the
        // (warp-uniform) branch that handles a new maximum
        if (s > m) {
          const float corr = __expf(m - s);                           // exp(-inf) = 0 on the first frame
          denom = fmaf(denom, corr, 1.f);
#pragma unroll
          for (int q = 0; q < 8; ++q) pooled[q] = fmaf(pooled[q], corr, x[q]);
          m = s;
        } else {
          const float e = __expf(s - m);
          denom += e;
#pragma unroll
          for (int q = 0; q < 8; ++q) pooled[q] = fmaf(e, x[q], pooled[q]);
        }
      }
      // the deferred 1 / (1 - p) of the frame_mlp dropout; average: clamp_min(count, 1) (:280); max of nothing but
      // invalid frames stays finfo.min
      const float inv = POOL == 2 ? ds.scale : ds.scale / (POOL == 1 ? fmaxf(denom, 1.f) : denom);
#pragma unroll
      for (int q = 0; q < 8; ++q) pooled[q] = (POOL == 2 && pooled[q] == -FLT_MAX) ? pooled[q] : pooled[q] * inv;
      if (RAW) {
        *reinterpret_cast<float4*>(z + (size_t)row * FH + c0) = make_float4(pooled[0], pooled[1], pooled[2], pooled[3]);
        *reinterpret_cast<float4*>(z + (size_t)row * FH + c1) = make_float4(pooled[4], pooled[5], pooled[6], pooled[7]);
        continue;
      }
      if (drop) {     // site FRAME_POOL (32-bit draws): element e = col, block e >> 2
        const Philox4 r0 = philox4x32_10_rk((uint32_t)(2 * lane), g, pass, AVHOT_SITE_FRAME_POOL, ds);
        const Philox4 r1 = philox4x32_10_rk((uint32_t)(2 * lane + 1), g, pass, AVHOT_SITE_FRAME_POOL, ds);
        const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) pooled[q] = w[q] >= ds.thr ? pooled[q] * ds.scale : 0.f;
      }
      // LayerNorm(H), biased variance, eps 1e-5 (nn.LayerNorm defaults)
      float sum = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) sum += pooled[q];
      const float mean = warp_sum(sum) * (1.f / FH);
      float sq = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float d = pooled[q] - mean; sq = fmaf(d, d, sq); }
      const float rstd = 1.f / sqrtf(warp_sum(sq) * (1.f / FH) + 1e-5f);
      const float4 lw0 = __ldg(reinterpret_cast<const float4*>(ln_w + c0));
      const float4 lw1 = __ldg(reinterpret_cast<const float4*>(ln_w + c1));
      const float4 lb0 = __ldg(reinterpret_cast<const float4*>(ln_b + c0));
      const float4 lb1 = __ldg(reinterpret_cast<const float4*>(ln_b + c1));
      const float lw[8] = {lw0.x, lw0.y, lw0.z, lw0.w, lw1.x, lw1.y, lw1.z, lw1.w};
      const float lb[8] = {lb0.x, lb0.y, lb0.z, lb0.w, lb1.x, lb1.y, lb1.z, lb1.w};
#pragma unroll
      for (int q = 0; q < 8; ++q) y[q] = (pooled[q] - mean) * rstd * lw[q] + lb[q];
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) y[q] = 0.f;
    }
    if (!PROJECT) {
      if (row < n_rows) {
        *reinterpret_cast<float4*>(z + (size_t)row * FH + c0) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(z + (size_t)row * FH + c1) = make_float4(y[4], y[5], y[6], y[7]);
      }
      continue;
    }
    // staged [column][row]; the 16-byte row groups of a column are XOR-permuted by the owning lane (= column >> 3)
    // so the 32 lanes of this store do not all hit one bank (4-way instead of 32-way conflicts)
    const int ri = (((i >> 2) ^ (lane & 7)) << 2) | (i & 3);
#pragma unroll
    for (int q = 0; q < 4; ++q) { ys[c0 + q][ri] = y[q]; ys[c1 + q][ri] = y[4 + q]; }
  }
  if (!PROJECT) return;
  __syncthreads();
  // projection H -> out for the 32 rows of this CTA: thread = (output o, half of the rows)
  const int half = tid >> 7;
  for (int o = tid & 127; o < out_dim; o += 128) {
    float acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = 0.f;
#pragma unroll 4
    for (int j = 0; j < FH; ++j) {
      const float w = __ldg(proj_wt + (size_t)j * out_dim + o);
#pragma unroll
      for (int r4 = 0; r4 < 16; r4 += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&ys[j][(((half * 16 + r4) >> 2) ^ ((j >> 3) & 7)) << 2]);
        acc[r4] = fmaf(w, v.x, acc[r4]); acc[r4 + 1] = fmaf(w, v.y, acc[r4 + 1]);
        acc[r4 + 2] = fmaf(w, v.z, acc[r4 + 2]); acc[r4 + 3] = fmaf(w, v.w, acc[r4 + 3]);
      }
    }
    const float pb = __ldg(proj_b + o);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const long long row = row0 + half * 16 + r;
      if (row < n_rows) z[(size_t)row * out_dim + o] = acc[r] + pb;
    }
  }
}

static size_t align256f(size_t x) { return (x + 255) & ~(size_t)255; }

// u8 pixels -> fp32 k / 255 (src/dataprocessing.py:261-265), 16 pixels per thread, for the tensor-core GEMM
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, size_t n16) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n16) return;
  const uint4 p = __ldg(reinterpret_cast<const uint4*>(in) + i);
  const uint32_t w[4] = {p.x, p.y, p.z, p.w};
  float4* o = reinterpret_cast<float4*>(out) + i * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = make_float4(__fdiv_rn((float)(w[j] & 255u), 255.f), __fdiv_rn((float)((w[j] >> 8) & 255u), 255.f),
                       __fdiv_rn((float)((w[j] >> 16) & 255u), 255.f), __fdiv_rn((float)(w[j] >> 24), 255.f));
}

}  // namespace avhot

using namespace avhot;

extern "C" int avhot_frame_create(const avhot_frame_desc* d, avhot_frame_t* out, avhot_stream_t stream) {
  AVHOT_REQUIRE(d && out, AVHOT_ERR_INVALID, "frame_create: null argument");
  AVHOT_REQUIRE(d->frame_dim > 0 && d->output_dim > 0, AVHOT_ERR_INVALID, "frame_create: bad dims");
  AVHOT_REQUIRE(d->hidden_dim == FH, AVHOT_ERR_UNSUPPORTED,
                "frame_create: hidden_dim=%d; this build is specialised for hidden_dim=256 "
                "(configs/base.yaml:44)", d->hidden_dim);
  AVHOT_REQUIRE((d->frame_dim & 3) == 0, AVHOT_ERR_UNSUPPORTED, "frame_create: frame_dim=%d must be a multiple of 4",
                d->frame_dim);
  // attention.* may be NULL: FrameEncoder(temporal_pooling='average' | 'max') has no attention layer (src/encoders.py:236-241)
  AVHOT_REQUIRE(d->mlp_weight && d->mlp_bias && (d->attn_weight != nullptr) == (d->attn_bias != nullptr) && d->ln_weight &&
                    d->ln_bias && d->proj_weight && d->proj_bias, AVHOT_ERR_INVALID, "frame_create: tensors missing");
  const int H = FH, O = d->output_dim;
  const size_t sizes[8] = {(size_t)H * d->frame_dim, (size_t)H, (size_t)H, 1, (size_t)H, (size_t)H, (size_t)O * H, (size_t)O};
  size_t total = 0;
  for (size_t n : sizes) total += align256f(n * sizeof(float));
  const bool u8_tc = (d->frame_dim % 128) == 0;           // uint8 frames straight onto the tensor cores (gemm_u8.cu)
  const size_t wq_bytes = u8_tc ? align256f((size_t)H * d->frame_dim * 2) + 512 : 0;
  total += wq_bytes;
  avhot_frame* h = new avhot_frame();
  h->frame_dim = d->frame_dim; h->hidden_dim = H; h->output_dim = O;
  h->arena = nullptr;
  cudaError_t e = cudaMalloc(&h->arena, total);
  if (e != cudaSuccess) {
    delete h;
    set_error("frame_create: cudaMalloc(%zu) -> %s", total, cudaGetErrorString(e));
    return AVHOT_ERR_CUDA;
  }
  cudaStream_t s = as_stream(stream);
  char* p = reinterpret_cast<char*>(h->arena);
  float* dst[8];
  for (int i = 0; i < 8; ++i) { dst[i] = reinterpret_cast<float*>(p); p += align256f(sizes[i] * sizeof(float)); }
  h->mlp_w = dst[0]; h->mlp_b = dst[1]; h->attn_w = dst[2]; h->attn_b = dst[3];
  h->ln_w = dst[4]; h->ln_b = dst[5]; h->proj_wt = dst[6]; h->proj_b = dst[7];
  const float* src[8] = {d->mlp_weight, d->mlp_bias, d->attn_weight, d->attn_bias, d->ln_weight, d->ln_bias,
                         nullptr, d->proj_bias};
  int rc = AVHOT_OK;
  h->has_attention = d->attn_weight != nullptr;
  if (!h->has_attention && cudaMemsetAsync(dst[2], 0, align256f(sizes[2] * sizeof(float)) + align256f(sizeof(float)), s) != cudaSuccess)
    rc = AVHOT_ERR_CUDA;
  for (int i = 0; i < 8 && rc == AVHOT_OK; ++i) {
    if (!src[i]) continue;
    if (cudaMemcpyAsync(dst[i], src[i], sizes[i] * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
      set_error("frame_create: device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = AVHOT_ERR_CUDA;
    }
  }
  if (rc == AVHOT_OK) rc = transpose(d->proj_weight, h->proj_wt, O, H, s);
  h->wq_f16 = nullptr;
  h->wq_inv = nullptr;
  if (rc == AVHOT_OK && u8_tc) {
    h->wq_f16 = p;
    h->wq_inv = reinterpret_cast<float*>(p + align256f((size_t)H * d->frame_dim * 2));
    rc = gemm_u8_pack_weights(d->mlp_weight, H, d->frame_dim, h->wq_f16, h->wq_inv,
                              reinterpret_cast<uint32_t*>(h->wq_inv) + 64, s);
  }
  if (rc != AVHOT_OK) {
    cudaFree(h->arena);
    delete h;
    return rc;
  }
  *out = h;
  return AVHOT_OK;
}

extern "C" void avhot_frame_destroy(avhot_frame_t h) {
  if (!h) return;
  cudaFree(h->arena);
  delete h;
}

// scratch: only the tensor-core path with u8 frames needs one (the fp32 image of the pixels)
extern "C" size_t avhot_frame_workspace_bytes(avhot_frame_t h, int B, int F, int precision) {
  if (!h || B <= 0 || F <= 0 || precision != AVHOT_PREC_TC) return 0;
  return sizeof(float) * (size_t)B * F * h->frame_dim;
}

extern "C" int avhot_frame_hidden(avhot_frame_t h, const float* frames_f32, const uint8_t* frames_u8, float* hid,
                                  int B, int F, int precision, void* workspace, size_t workspace_bytes,
                                  avhot_stream_t stream) {
  AVHOT_REQUIRE(h && hid && B >= 0 && F >= 1, AVHOT_ERR_INVALID, "frame_hidden: bad arguments");
  AVHOT_REQUIRE((frames_f32 != nullptr) != (frames_u8 != nullptr), AVHOT_ERR_INVALID,
                "frame_hidden: exactly one of frames_f32 / frames_u8");
  AVHOT_REQUIRE(precision == AVHOT_PREC_FP32 || precision == AVHOT_PREC_TC, AVHOT_ERR_UNSUPPORTED,
                "frame_hidden: precision %d not built", precision);
  AVHOT_REQUIRE((long long)B * F <= 0x7fffffff / 2, AVHOT_ERR_INVALID, "frame_hidden: B*F too large for one call");
  if (precision == AVHOT_PREC_TC && frames_u8 && h->wq_f16 &&
      gemm_u8_pair_tc_supported(frames_u8, h->frame_dim, h->hidden_dim, h->frame_dim))
    return gemm_u8_pair_tc(frames_u8, h->frame_dim, h->wq_f16, h->mlp_b, h->wq_inv, hid, h->hidden_dim, B * F, h->hidden_dim,
                           h->frame_dim, ACT_RELU, as_stream(stream), "frame_gemm");
  if (precision == AVHOT_PREC_TC && frames_u8 && h->wq_f16 && gemm_u8_tc_supported(frames_u8, h->frame_dim, h->hidden_dim, h->frame_dim))
    return gemm_u8_tc(frames_u8, h->frame_dim, h->wq_f16, h->mlp_b, h->wq_inv, hid, h->hidden_dim, B * F, h->hidden_dim,
                      h->frame_dim, ACT_RELU, as_stream(stream), "frame_gemm");
  if (precision == AVHOT_PREC_TC && frames_u8 && (((size_t)B * F * h->frame_dim) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(frames_u8) & 15) == 0 && workspace &&
      workspace_bytes >= avhot_frame_workspace_bytes(h, B, F, precision)) {
    // u8 ingest on the tensor-core path: widen to fp32 k/255 once, then the same TF32 GEMM
    const size_t n16 = (size_t)B * F * h->frame_dim / 16;
    {
      ProfScope prof("frame_u8_to_f32", as_stream(stream));
      u8_to_f32_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, as_stream(stream)>>>(
          frames_u8, reinterpret_cast<float*>(workspace), n16);
      AVHOT_LAUNCH_CHECK();
    }
    frames_f32 = reinterpret_cast<const float*>(workspace);
    frames_u8 = nullptr;
  }
  // fp32 frames on the tensor-core path: converted to fp16 inside a CTA-pair GEMM that shares the fp16 weights
  // (gemm_cvt.cu; wq_f16 is W s / 255, so the pixel scaling is undone by 255); TF32 on the fp32 operands otherwise
  if (precision == AVHOT_PREC_TC && frames_f32 && h->wq_f16 &&
      gemm_f32cvt_tc_supported(frames_f32, h->frame_dim, h->hidden_dim, h->frame_dim))
    return gemm_f32cvt_tc(frames_f32, h->frame_dim, h->wq_f16, h->mlp_b, h->wq_inv, 255.f, hid, h->hidden_dim, B * F,
                          h->hidden_dim, h->frame_dim, ACT_RELU, as_stream(stream), "frame_gemm");
  if (precision == AVHOT_PREC_TC && frames_f32)
    return gemm_tf32_tc(frames_f32, h->frame_dim, h->mlp_w, h->frame_dim, h->mlp_b, hid, h->hidden_dim, B * F,
                        h->hidden_dim, h->frame_dim, ACT_RELU, as_stream(stream), "frame_gemm");
  return gemm_f32(frames_f32, frames_u8, h->frame_dim, h->mlp_w, h->frame_dim, h->mlp_b, hid, h->hidden_dim,
                  B * F, h->hidden_dim, h->frame_dim, ACT_RELU, as_stream(stream), "frame_gemm_f32");
}

static int frame_tail_impl(avhot_frame_t h, const float* hid, const uint8_t* frame_valid, float* z, int B, int F, int T,
                           float p, uint64_t seed, uint64_t clip0, const uint32_t* clip_ids, uint32_t t0, bool project,
                           avhot_stream_t stream, int pooling = AVHOT_POOL_ATTENTION, bool raw = false) {
  AVHOT_REQUIRE(h && hid && z && B >= 0 && F >= 1 && T >= 1, AVHOT_ERR_INVALID, "frame_tail: bad arguments");
  AVHOT_REQUIRE(p >= 0.f && p < 1.f, AVHOT_ERR_INVALID, "frame_tail: dropout p=%g outside [0, 1)", (double)p);
  AVHOT_REQUIRE((long long)F * 64 + 64 <= 0xffffffffLL, AVHOT_ERR_INVALID, "frame_tail: F too large");
  AVHOT_REQUIRE(h->has_attention || (pooling != AVHOT_POOL_ATTENTION && !raw), AVHOT_ERR_INVALID,
                "frame_tail: attention pooling on an encoder created without attention.weight / attention.bias");
  AVHOT_REQUIRE(project || (reinterpret_cast<uintptr_t>(z) & 15) == 0, AVHOT_ERR_INVALID,
                "frame_tail_norm: y must be 16-byte aligned");
  if (B == 0) return AVHOT_OK;
  const long long rows = (long long)B * T;
  const long long grid = (rows + FT_ROWS - 1) / FT_ROWS;
  AVHOT_REQUIRE(rows < 0x7fffffffLL, AVHOT_ERR_INVALID, "frame_tail: B*T too large for one call");
  DropSpec ds = make_drop(p, seed, clip0, t0);
  ds.clip_ids = clip_ids;
  ProfScope prof("frame_tail", as_stream(stream));
#define AVHOT_TAIL_LAUNCH(...)                                                                                       \
  frame_tail_kernel<__VA_ARGS__><<<(unsigned)grid, FT_WARPS * 32, 0, as_stream(stream)>>>(                            \
      hid, frame_valid, z, h->attn_w, h->attn_b, h->ln_w, h->ln_b, h->proj_wt, h->proj_b, B, F, T, h->output_dim, ds)
  if (raw) {
    AVHOT_TAIL_LAUNCH(false, 0, true);
  } else if (pooling == AVHOT_POOL_AVERAGE) {
    AVHOT_TAIL_LAUNCH(true, 1, false);
  } else if (pooling == AVHOT_POOL_MAX) {
    AVHOT_TAIL_LAUNCH(true, 2, false);
  } else if (project)
    frame_tail_kernel<true><<<(unsigned)grid, FT_WARPS * 32, 0, as_stream(stream)>>>(
        hid, frame_valid, z, h->attn_w, h->attn_b, h->ln_w, h->ln_b, h->proj_wt, h->proj_b, B, F, T, h->output_dim, ds);
  else
    frame_tail_kernel<false><<<(unsigned)grid, FT_WARPS * 32, 0, as_stream(stream)>>>(
        hid, frame_valid, z, h->attn_w, h->attn_b, h->ln_w, h->ln_b, h->proj_wt, h->proj_b, B, F, T, h->output_dim, ds);
#undef AVHOT_TAIL_LAUNCH
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

extern "C" int avhot_frame_tail_pool(avhot_frame_t h, const float* hid, const uint8_t* frame_valid, float* z, int B,
                                     int F, int T, float p, uint64_t seed, uint64_t clip0, uint32_t t0, int pooling,
                                     avhot_stream_t stream) {
  AVHOT_REQUIRE(pooling == AVHOT_POOL_ATTENTION || pooling == AVHOT_POOL_AVERAGE || pooling == AVHOT_POOL_MAX,
                AVHOT_ERR_INVALID, "frame_tail_pool: unknown pooling %d", pooling);
  return frame_tail_impl(h, hid, frame_valid, z, B, F, T, p, seed, clip0, nullptr, t0, true, stream, pooling);
}

extern "C" int avhot_frame_attention_pool(avhot_frame_t h, const float* frames, const uint8_t* frame_valid,
                                          float* pooled, int B, int F, avhot_stream_t stream) {
  return frame_tail_impl(h, frames, frame_valid, pooled, B, F, 1, 0.f, 0, 0, nullptr, 0, false, stream,
                         AVHOT_POOL_ATTENTION, true);
}

extern "C" int avhot_frame_tail(avhot_frame_t h, const float* hid, const uint8_t* frame_valid, float* z, int B,
                                int F, int T, float p, uint64_t seed, uint64_t clip0, uint32_t t0,
                                avhot_stream_t stream) {
  return frame_tail_impl(h, hid, frame_valid, z, B, F, T, p, seed, clip0, nullptr, t0, true, stream);
}

extern "C" int avhot_frame_tail_ids(avhot_frame_t h, const float* hid, const uint8_t* frame_valid, float* z, int B,
                                    int F, int T, float p, uint64_t seed, const uint32_t* clip_ids, uint32_t t0,
                                    avhot_stream_t stream) {
  AVHOT_REQUIRE(clip_ids, AVHOT_ERR_INVALID, "frame_tail_ids: clip_ids is NULL");
  return frame_tail_impl(h, hid, frame_valid, z, B, F, T, p, seed, 0, clip_ids, t0, true, stream);
}

extern "C" int avhot_frame_tail_norm(avhot_frame_t h, const float* hid, const uint8_t* frame_valid, float* y, int B,
                                     int F, int T, float p, uint64_t seed, uint64_t clip0, uint32_t t0,
                                     avhot_stream_t stream) {
  return frame_tail_impl(h, hid, frame_valid, y, B, F, T, p, seed, clip0, nullptr, t0, false, stream);
}
