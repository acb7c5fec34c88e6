// SequenceEncoder(encoder_type='gru')  (src/encoders.py:65-77 ctor, :137-155 forward).
//
//   seq (B, T, in) -> nn.GRU(in -> H, L layers, batch_first, zero initial state) -> h_n[-1] -> Linear(H -> out)
//
// nn.GRU, gate order r, z, n (weight rows [0,H) r, [H,2H) z, [2H,3H) n):
//   r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)      z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
//   n = tanh(W_in x + b_in + r * (W_hn h + b_hn))   h' = (1 - z) * n + z * h
// fp32 CUDA-core path in the style of lstm.cu: per layer the input projections xp = x W_ih^T + b_ih for all steps at
// once (gemm.cu), then the recurrence - one CTA per tile of BT clips, thread j owns hidden unit j, the three hidden
// projections of unit j for every clip of the tile live in registers, h_{t-1} of the tile is broadcast from shared
// memory, W_hh is read k-major so every warp load is one coalesced 128-byte line.  Inter-layer dropout is inactive
// in eval mode and under MC-dropout.
#include "common.cuh"

struct avhot_gru {
  int input_dim, hidden_dim, num_layers, output_dim;
  float* w_ih[4];    // (3H, in_k) row-major copies
  float* w_hh_t[4];  // [k][gate][j] = W_hh[gate*H + j][k]
  float* b_ih[4];    // (3H)
  float* b_hh[4];    // (3H)
  float* proj_w;     // (out, H)
  float* proj_b;
  float* arena;
};

namespace avhot {

constexpr int GH = 256;   // hidden size this kernel is specialised for (one thread per unit)

__device__ __forceinline__ float gru_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

template <int BT>
__global__ void __launch_bounds__(GH)
gru_recurrence_kernel(const float* __restrict__ xp,         // (B, T, 3H), includes b_ih
                      const float* __restrict__ whh_t,      // [H][3][H]
                      const float* __restrict__ b_hh,       // (3H)
                      float* __restrict__ h_seq,            // (B, T, H) or NULL
                      float* __restrict__ h_last,           // (B, H) or NULL
                      const int32_t* __restrict__ lengths,  // (B) valid steps per clip, or NULL (= T)
                      int B, int T) {
  constexpr int HP = BT + 4;
  extern __shared__ __align__(16) unsigned char gru_smem_raw[];
  float (*hs)[GH][HP] = reinterpret_cast<float (*)[GH][HP]>(gru_smem_raw);   // [2][GH][HP]
  const int j = threadIdx.x;
  const int b0 = blockIdx.x * BT;
#pragma unroll
  for (int b = 0; b < BT; ++b) hs[0][j][b] = 0.f;
  const float bh[3] = {__ldg(b_hh + j), __ldg(b_hh + GH + j), __ldg(b_hh + 2 * GH + j)};
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const int cur = t & 1;
    float acc[3][BT];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[g][b] = bh[g];
    const float* w = whh_t + j;
#pragma unroll 2
    for (int k = 0; k < GH; ++k) {
      float wg[3];
#pragma unroll
      for (int g = 0; g < 3; ++g) wg[g] = __ldg(w + (size_t)(k * 3 + g) * GH);
      float hv[BT];
#pragma unroll
      for (int b4 = 0; b4 < BT; b4 += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&hs[cur][k][b4]);
        hv[b4] = v.x; hv[b4 + 1] = v.y; hv[b4 + 2] = v.z; hv[b4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int b = 0; b < BT; ++b) acc[g][b] = fmaf(wg[g], hv[b], acc[g][b]);
    }
    float hn[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const bool live = b0 + b < B;
      const float* x = xp + ((size_t)(live ? b0 + b : b0) * T + t) * (3 * GH) + j;
      const float r = gru_sigmoid(__ldg(x) + acc[0][b]);
      const float z = gru_sigmoid(__ldg(x + GH) + acc[1][b]);
      const float n = tanhf(fmaf(r, acc[2][b], __ldg(x + 2 * GH)));
      hn[b] = fmaf(z, hs[cur][j][b] - n, n);           // (1 - z) n + z h
    }
#pragma unroll
    for (int b4 = 0; b4 < BT; b4 += 4)
      *reinterpret_cast<float4*>(&hs[cur ^ 1][j][b4]) = make_float4(hn[b4], hn[b4 + 1], hn[b4 + 2], hn[b4 + 3]);
    if (h_seq) {
#pragma unroll
      for (int b = 0; b < BT; ++b)
        if (b0 + b < B) h_seq[((size_t)(b0 + b) * T + t) * GH + j] = hn[b];
    }
    if (h_last) {
#pragma unroll
      for (int b = 0; b < BT; ++b)
        if (b0 + b < B) {   // h_n = state after the clip's last valid step (packed sequences, src/encoders.py:139-147)
          const int last_t = (lengths ? min(max(__ldg(lengths + b0 + b), 1), T) : T) - 1;
          if (t == last_t) h_last[(size_t)(b0 + b) * GH + j] = hn[b];
        }
    }
    __syncthreads();
  }
}

static size_t galign(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace avhot

using namespace avhot;

extern "C" int avhot_gru_create(const avhot_lstm_desc* d, avhot_gru_t* out, avhot_stream_t stream) {
  AVHOT_REQUIRE(d && out, AVHOT_ERR_INVALID, "gru_create: null argument");
  AVHOT_REQUIRE(d->num_layers >= 1 && d->num_layers <= 4, AVHOT_ERR_INVALID, "gru_create: num_layers=%d (1..4)",
                d->num_layers);
  AVHOT_REQUIRE(d->input_dim > 0 && d->output_dim > 0, AVHOT_ERR_INVALID, "gru_create: bad dims");
  AVHOT_REQUIRE(d->hidden_dim == GH, AVHOT_ERR_UNSUPPORTED,
                "gru_create: hidden_dim=%d; this build is specialised for hidden_dim=256", d->hidden_dim);
  AVHOT_REQUIRE((d->input_dim & 3) == 0, AVHOT_ERR_UNSUPPORTED, "gru_create: input_dim=%d must be a multiple of 4",
                d->input_dim);
  const int H = GH, L = d->num_layers;
  size_t total = 0;
  for (int l = 0; l < L; ++l) {
    const int in = l == 0 ? d->input_dim : H;
    AVHOT_REQUIRE(d->weight_ih[l] && d->weight_hh[l] && d->bias_ih[l] && d->bias_hh[l], AVHOT_ERR_INVALID,
                  "gru_create: layer %d tensors missing", l);
    total += galign(sizeof(float) * 3 * H * in) + galign(sizeof(float) * 3 * H * H) + 2 * galign(sizeof(float) * 3 * H);
  }
  AVHOT_REQUIRE(d->proj_weight && d->proj_bias, AVHOT_ERR_INVALID, "gru_create: projection missing");
  total += galign(sizeof(float) * d->output_dim * H) + galign(sizeof(float) * d->output_dim);
  avhot_gru* h = new avhot_gru();
  h->input_dim = d->input_dim; h->hidden_dim = H; h->num_layers = L; h->output_dim = d->output_dim;
  h->arena = nullptr;
  cudaError_t e = cudaMalloc(&h->arena, total);
  if (e != cudaSuccess) {
    delete h;
    set_error("gru_create: cudaMalloc(%zu) -> %s", total, cudaGetErrorString(e));
    return AVHOT_ERR_CUDA;
  }
  cudaStream_t s = as_stream(stream);
  char* p = reinterpret_cast<char*>(h->arena);
  auto take = [&](size_t bytes) { float* r = reinterpret_cast<float*>(p); p += galign(bytes); return r; };
  auto copy = [&](float* dst, const float* src, size_t n) {
    return cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s) == cudaSuccess;
  };
  bool ok = true;
  int rc = AVHOT_OK;
  for (int l = 0; l < L && ok && rc == AVHOT_OK; ++l) {
    const int in = l == 0 ? d->input_dim : H;
    h->w_ih[l] = take(sizeof(float) * 3 * H * in);
    h->w_hh_t[l] = take(sizeof(float) * 3 * H * H);
    h->b_ih[l] = take(sizeof(float) * 3 * H);
    h->b_hh[l] = take(sizeof(float) * 3 * H);
    ok = copy(h->w_ih[l], d->weight_ih[l], (size_t)3 * H * in) && copy(h->b_ih[l], d->bias_ih[l], 3 * H) &&
         copy(h->b_hh[l], d->bias_hh[l], 3 * H);
    if (ok) rc = transpose(d->weight_hh[l], h->w_hh_t[l], 3 * H, H, s);   // [k][gate*H + j]
  }
  if (ok && rc == AVHOT_OK) {
    h->proj_w = take(sizeof(float) * d->output_dim * H);
    h->proj_b = take(sizeof(float) * d->output_dim);
    ok = copy(h->proj_w, d->proj_weight, (size_t)d->output_dim * H) && copy(h->proj_b, d->proj_bias, d->output_dim);
  }
  if (!ok || rc != AVHOT_OK) {
    if (!ok) { set_error("gru_create: device copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = AVHOT_ERR_CUDA; }
    cudaFree(h->arena);
    delete h;
    return rc;
  }
  *out = h;
  return AVHOT_OK;
}

extern "C" void avhot_gru_destroy(avhot_gru_t h) {
  if (!h) return;
  cudaFree(h->arena);
  delete h;
}

// workspace: xp (B*T*3H) + h_seq (B*T*H) + h_last (B*H), fp32
extern "C" size_t avhot_gru_workspace_bytes(avhot_gru_t h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return 0;
  const size_t H = h->hidden_dim;
  return galign(sizeof(float) * (size_t)B * T * 3 * H) + galign(sizeof(float) * (size_t)B * T * H) +
         galign(sizeof(float) * (size_t)B * H);
}

extern "C" int avhot_gru_forward(avhot_gru_t h, const float* seq, const int32_t* lengths, float* out, int B, int T,
                                 void* workspace, size_t workspace_bytes, avhot_stream_t stream) {
  AVHOT_REQUIRE(h && seq && out && B >= 0 && T >= 1, AVHOT_ERR_INVALID, "gru_forward: bad arguments");
  if (B == 0) return AVHOT_OK;
  const size_t need = avhot_gru_workspace_bytes(h, B, T);
  AVHOT_REQUIRE(workspace && workspace_bytes >= need, AVHOT_ERR_WORKSPACE, "gru_forward: workspace %zu < %zu bytes",
                workspace_bytes, need);
  AVHOT_REQUIRE((long long)B * T <= 0x7fffffff / 8, AVHOT_ERR_INVALID, "gru_forward: B*T too large for one call");
  const int H = h->hidden_dim;
  cudaStream_t s = as_stream(stream);
  char* p = reinterpret_cast<char*>(workspace);
  float* xp = reinterpret_cast<float*>(p); p += galign(sizeof(float) * (size_t)B * T * 3 * H);
  float* h_seq = reinterpret_cast<float*>(p); p += galign(sizeof(float) * (size_t)B * T * H);
  float* h_last = reinterpret_cast<float*>(p);
  const bool small = ceil_div(B, 32) * 2 < num_sms();   // few clips: smaller tiles, more CTAs
  for (int l = 0; l < h->num_layers; ++l) {
    const float* x = l == 0 ? seq : h_seq;
    const int in = l == 0 ? h->input_dim : H;
    const bool last = l == h->num_layers - 1;
    int rc = gemm_f32(x, nullptr, in, h->w_ih[l], in, h->b_ih[l], xp, 3 * H, B * T, 3 * H, in, ACT_NONE, s,
                      "gru_input_gemm");
    if (rc != AVHOT_OK) return rc;
    ProfScope prof("gru_recurrence", s);
    if (small) {
      constexpr int BT = 4;
      const size_t smem = sizeof(float) * 2 * GH * (BT + 4);
      gru_recurrence_kernel<BT><<<ceil_div(B, BT), GH, smem, s>>>(xp, h->w_hh_t[l], h->b_hh[l], last ? nullptr : h_seq,
                                                                  last ? h_last : nullptr, lengths, B, T);
    } else {
      constexpr int BT = 32;
      const size_t smem = sizeof(float) * 2 * GH * (BT + 4);
      AVHOT_CUDA(cudaFuncSetAttribute(gru_recurrence_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      gru_recurrence_kernel<BT><<<ceil_div(B, BT), GH, smem, s>>>(xp, h->w_hh_t[l], h->b_hh[l], last ? nullptr : h_seq,
                                                                  last ? h_last : nullptr, lengths, B, T);
    }
    AVHOT_LAUNCH_CHECK();
  }
  return gemm_f32(h_last, nullptr, H, h->proj_w, H, h->proj_b, out, h->output_dim, B, h->output_dim, H, ACT_NONE, s,
                  "gru_projection");
}
