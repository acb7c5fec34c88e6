// a3: SequenceEncoder(encoder_type='lstm')  (src/encoders.py:49-63 ctor, :137-155 forward).
//
//   seq (B, T, in) -> nn.LSTM(in -> H, L layers, batch_first, zero initial state) -> h_n[-1]
//       -> Linear(H -> out)                                             = (B, out)
//
// AVHOT_PREC_FP32 path (this file): per layer
//   1. xp = x @ W_ih^T + (b_ih + b_hh) for all steps at once (gemm.cu), (B*T, 4H);
//   2. the recurrence, one CTA per tile of BT clips, thread j owns hidden unit j: the four gate
//      pre-activations of unit j for every clip of the tile live in registers, h_{t-1} of the
//      tile is broadcast from shared memory, W_hh is read k-major ([k][gate][j], transposed once
//      at create time) so each warp load is one coalesced 128-byte line served from L1/L2;
//      c_t stays in registers for the whole sequence.
// Gate order i, f, g, o (nn.LSTM); c' = f*c + i*g; h' = o*tanh(c').  Inter-layer dropout is
// inactive in eval mode and under MC-dropout.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

struct avhot_lstm {
  int input_dim, hidden_dim, num_layers, output_dim;
  float* w_ih[4];    // (4H, in_k) row-major copies
  float* w_hh_t[4];  // [k][j][gate]  = W_hh[gate*H + j][k]
  float* bias[4];    // b_ih + b_hh
  float* bias_tc[4]; // the same with the i, f, o rows halved (tensor-core path: sigmoid(x) = 0.5 tanh(x / 2) + 0.5)
  void* w_hh_r16[4]; // fp16, rows regrouped per 32-unit chunk (lstm_tc.cu)
  void* w_ih_f16[4];  // fp16 input weights for the tensor-core path: layer 0 (4H, 3*in) = [W_hi | W_hi | W_lo],
                      // layers > 0 (4H, H)
  void* w_ih_p16[4];  // layers > 0: fp16 W_ih / W_hh / fp32 bias in the fused layer kernel's row order (lstm_fused.cu)
  void* w_hh_p16[4];
  float* bias_p[4];
  void* w_ih_fold_p16; // layer 0, (4H, 192) = [W'_hi | W'_hi | W'_lo], W' = W[:, :64] + W[:, 64:104] D: the MFCC columns of the
                       // front-end's operand are D x its log-mel columns, so their weights fold into the log-mel ones
  float* proj_w;     // (out, H)
  float* proj_b;
  float* arena;
};

namespace avhot {

constexpr int LH = 256;  // hidden size this kernel is specialised for (one thread per unit)

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

// W_hh is streamed through shared memory: chunks of KC k-rows (KC x 256 units x 4 gates; 64 KB at KC = 16, contiguous in
// the [k][j][gate] copy) travel global/L2 -> shared with cp.async into a ring of NB buffers, NB - 1 chunks ahead of the
// FMAs, one block barrier per chunk.  (Reading W with per-thread __ldg left two warps per scheduler waiting on L2 round
// trips: 111 k cycles per step measured at BT = 12 against 25 k of FMAs, whatever the unroll depth.)
//
// Thread = hidden unit j, all BT clips of the tile: per k-row one 128-bit load of the unit's four gate weights, BT / 4
// broadcast loads of h_{t-1}[k], and 4 x BT / 2 FFMA2 (fma.rn.f32x2 over clip pairs: each half rounds exactly like fmaf).
// The shape was settled with tools/micro/rec_loop.cu, which runs this loop in isolation: 512 threads with the clips split
// in halves read 96 shared-memory wavefronts per row - as many cycles as the 96 of the FFMA2s - and reached 117 cycles
// per row, 144 with a barrier every 8 rows, 191 in the real kernel; one thread per unit halves the wavefronts (56) and,
// with the next row's operands fetched into a second register set BEFORE this row's FMAs (ptxas otherwise sinks each
// load to within a few instructions of its use, and two warps per scheduler cannot cover that), reaches 107 cycles per
// row, 112 with a barrier every 16 rows.  Measured in the kernel: 15.5 -> 13.5 ms for the two layers of a 1 565-clip
// flush (166 cycles per row: the weight stream, the step's activations and xp loads are on top of the isolated loop).
// Tried on top and dropped: a producer warp feeding the ring with TMA bulk copies, multicast to a cluster of two CTAs
// (half the L2 traffic) and mbarrier hand-shakes instead of the block barrier - bit-identical, slower with or without
// the multicast, so the kernel is not L2-bound at this rate either; fetching xp a step ahead - no change.
constexpr int lstm_chunk_rows(int BT) { return BT <= 12 ? 16 : 8; }
constexpr int lstm_ring_depth(int BT) {
  const int room = (227 * 1024 - 1024 - (int)sizeof(float) * 2 * LH * (BT + 4)) / ((int)sizeof(float) * lstm_chunk_rows(BT) * 4 * LH);
  return room > 6 ? 6 : room;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int BT, int KC = lstm_chunk_rows(BT), int NB = lstm_ring_depth(BT)>
__global__ void __launch_bounds__(LH, 1)
lstm_recurrence_kernel(const float* __restrict__ xp,      // (B, T, 4H)
                       const float* __restrict__ whh_t,   // [H][H][4]: [k][j][gate] = W_hh[gate*H + j][k]
                       float* __restrict__ h_seq,         // (B, T, H) or NULL
                       float* __restrict__ h_last,        // (B, H) or NULL
                       const int32_t* __restrict__ lengths,  // (B) valid steps per clip, or NULL (= T)
                       int B, int T) {
  static_assert(BT % 4 == 0 && NB >= 3, "clip quads; ring: the chunk in use, one landing, one being requested");
  constexpr int HP = BT + 4;  // padded row: conflict-free 128-bit stores of one unit's BT values
  constexpr int NCH = LH / KC;
  extern __shared__ __align__(16) unsigned char lstm_smem_raw[];
  float (*hs)[LH][HP] = reinterpret_cast<float (*)[LH][HP]>(lstm_smem_raw);   // [2][LH][HP]
  float4 (*wb)[KC][LH] = reinterpret_cast<float4 (*)[KC][LH]>(lstm_smem_raw + sizeof(float) * 2 * LH * HP);  // [NB]
  const int j = threadIdx.x;
  const int b0 = blockIdx.x * BT;
  // chunk n of the endless stream (n = t * NCH + c) -> ring buffer n % NB; every thread copies KC x 16 bytes
  auto issue = [&](int n) {
    const float* src = whh_t + (size_t)(n % NCH) * (KC * 4 * LH);
    float* dst = reinterpret_cast<float*>(&wb[n % NB][0][0]);
#pragma unroll
    for (int i = 0; i < KC; ++i) cp_async16(dst + 4 * (j + i * LH), src + 4 * (j + i * LH));
    cp_async_commit();
  };
  const int n_total = T * NCH;
#pragma unroll
  for (int i = 0; i < NB - 1; ++i) {
    if (i < n_total) issue(i); else cp_async_commit();
  }
  float c[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    c[b] = 0.f;
    hs[0][j][b] = 0.f;
  }
  __syncthreads();
  int n = 0;
  for (int t = 0; t < T; ++t) {
    const int cur = t & 1;
    float2 acc[4][BT / 2];
#pragma unroll
    for (int b = 0; b < BT; b += 2) {
      const float* x0 = xp + ((size_t)min(b0 + b, B - 1) * T + t) * (4 * LH) + j;
      const float* x1 = xp + ((size_t)min(b0 + b + 1, B - 1) * T + t) * (4 * LH) + j;
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[g][b / 2] = make_float2(__ldg(x0 + g * LH), __ldg(x1 + g * LH));
    }
#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch, ++n) {
      cp_async_wait<NB - 2>();            // this thread's part of chunk n has landed (NB - 2 later ones may be in flight)
      __syncthreads();                    // ... and everybody's; everybody is also done with chunk n - 1's buffer
      if (n + NB - 1 < n_total) issue(n + NB - 1); else cp_async_commit();   // (empty group: keeps the wait count uniform)
      const float4* w = wb[n % NB][0];
      const float (*hrow)[HP] = &hs[cur][ch * KC];
      float4 wq[2], hq[2][BT / 4];
      wq[0] = w[j];
#pragma unroll
      for (int q = 0; q < BT / 4; ++q) hq[0][q] = *reinterpret_cast<const float4*>(&hrow[0][4 * q]);
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        if (kk + 1 < KC) {                // the next row's operands, before this row's FMAs
          wq[(kk + 1) & 1] = w[(kk + 1) * LH + j];
#pragma unroll
          for (int q = 0; q < BT / 4; ++q) hq[(kk + 1) & 1][q] = *reinterpret_cast<const float4*>(&hrow[kk + 1][4 * q]);
        }
        const float4 w4 = wq[kk & 1];
        const float2 wg[4] = {make_float2(w4.x, w4.x), make_float2(w4.y, w4.y), make_float2(w4.z, w4.z), make_float2(w4.w, w4.w)};
#pragma unroll
        for (int q = 0; q < BT / 4; ++q) {
          const float4 h4 = hq[kk & 1][q];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            acc[g][2 * q] = __ffma2_rn(wg[g], make_float2(h4.x, h4.y), acc[g][2 * q]);
            acc[g][2 * q + 1] = __ffma2_rn(wg[g], make_float2(h4.z, h4.w), acc[g][2 * q + 1]);
          }
        }
      }
    }
    float hn[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const float a0 = (b & 1) ? acc[0][b / 2].y : acc[0][b / 2].x, a1 = (b & 1) ? acc[1][b / 2].y : acc[1][b / 2].x;
      const float a2 = (b & 1) ? acc[2][b / 2].y : acc[2][b / 2].x, a3 = (b & 1) ? acc[3][b / 2].y : acc[3][b / 2].x;
      const float ig = sigmoid_acc(a0), fg = sigmoid_acc(a1);
      const float gg = tanhf(a2), og = sigmoid_acc(a3);
      c[b] = fmaf(fg, c[b], ig * gg);
      hn[b] = og * tanhf(c[b]);
    }
#pragma unroll
    for (int b4 = 0; b4 < BT; b4 += 4)
      *reinterpret_cast<float4*>(&hs[cur ^ 1][j][b4]) = make_float4(hn[b4], hn[b4 + 1], hn[b4 + 2], hn[b4 + 3]);
    if (h_seq) {
#pragma unroll
      for (int b = 0; b < BT; ++b)
        if (b0 + b < B) h_seq[((size_t)(b0 + b) * T + t) * LH + j] = hn[b];
    }
    if (h_last) {
#pragma unroll
      for (int b = 0; b < BT; ++b)
        if (b0 + b < B) {   // h_n = state after the clip's last valid step (packed sequences, src/encoders.py:139-147)
          const int last_t = (lengths ? min(max(__ldg(lengths + b0 + b), 1), T) : T) - 1;
          if (t == last_t) h_last[(size_t)(b0 + b) * LH + j] = hn[b];
        }
    }
    // no barrier here: the next step's first chunk barrier orders these h writes before anybody reads them
  }
  cp_async_wait<0>();
}

// W_hh (4H, H) -> [k][j][gate] = W_hh[gate*H + j][k]
__global__ void regroup_whh_kernel(const float* __restrict__ w_hh, float* __restrict__ dst) {
  const int k = blockIdx.x, j = threadIdx.x;
#pragma unroll
  for (int g = 0; g < 4; ++g) dst[((size_t)k * LH + j) * 4 + g] = w_hh[((size_t)g * LH + j) * LH + k];
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// fp16 operand split.  hi = fp16(x), lo = fp16(x - hi).  A' = [hi | lo | hi] against W' = [hi | hi | lo] turns one
// fp16 tensor-core GEMM over K' = 3K into A W^T - lo(A) lo(W)^T - O(2^-22), i.e. an fp32-class product: used for
// layer 0's input projection, whose inputs (log-mel features in dB, |x| ~ 100) are far too large for a single
// half-precision pass (measured 5.7e-3 error with one TF32 pass).  Requires |x| < 65504.
// Column layout (both operands): the features go in blocks of 64; block j (features 64 j .. 64 j + w - 1, w = 64 or the
// remainder) occupies columns 192 j .. 192 j + 3 w - 1 as [term 0 (w) | term 1 (w) | term 2 (w)], zeros up to the
// pitch.  For K <= 64 that is [hi | lo | hi]; for the 104 front-end features it is [lm_hi(64) | lm_lo(64) | lm_hi(64) |
// mf_hi(40) | mf_lo(40) | mf_hi(40) | 0(8)]: every 64-column k-block of the log-mel part is one whole term, which is what
// lets the front-end kernel hand its tensor-core operand tiles (hi, lo) to TMA as they are (frontend.cu).
__host__ __device__ __forceinline__ bool split3_slot(int c, int K, int& term, int& feat) {
  const int j = c / 192, off = c - 192 * j;
  const int w = min(64, K - 64 * j);
  if (w <= 0 || off >= 3 * w) return false;
  term = off / w;
  feat = 64 * j + off - term * w;
  return true;
}
// weights: W (N, K) fp32 -> out (N, ldo) fp16, term 0 and 1 = W_hi, term 2 = W_lo (the bias is added in fp32 by the GEMM
// epilogue).  Rows of the sigmoid gates (i, f, o: every gate but g = rows [2H, 3H)) are stored halved - exact, a power of
// two - so that the recurrent kernel's sigmoid(x) = 0.5 tanh(x / 2) + 0.5 needs no pre-scaling multiply (lstm_tc.cu).
__device__ __forceinline__ float tc_gate_scale(int n, int N) { return (n / (N / 4)) == 2 ? 1.f : 0.5f; }
// perm: output row p holds source row lstm_fused_src_row(p) (the fused layer kernel's gate-column order)
__global__ void pack_w_split3_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                     __half* __restrict__ out, int N, int K, int ldo, bool perm) {
  const int n = perm ? lstm_fused_src_row(blockIdx.x) : blockIdx.x;
  const float sc = tc_gate_scale(n, N);
  for (int k = threadIdx.x; k < ldo; k += blockDim.x) {
    int term, feat;
    __half o = __float2half_rn(0.f);
    if (split3_slot(k, K, term, feat)) {
      const float v = sc * w[(size_t)n * K + feat];
      const __half hi = __float2half_rn(v);
      o = term == 2 ? __float2half_rn(v - __half2float(hi)) : hi;
    }
    out[(size_t)blockIdx.x * ldo + k] = o;
  }
}
// weights of layers > 0: W (N, K) fp32 -> out (N, ldo) fp16 = [W | 0...]
__global__ void scale_gate_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int N, bool perm) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const int n = perm ? lstm_fused_src_row(p) : p;
  out[p] = tc_gate_scale(n, N) * b[n];
}
__global__ void pack_w_bias_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                   __half* __restrict__ out, int N, int K, int ldo, bool perm) {
  const int n = perm ? lstm_fused_src_row(blockIdx.x) : blockIdx.x;
  for (int k = threadIdx.x; k < ldo; k += blockDim.x) {
    float v = 0.f;
    if (k < K) v = tc_gate_scale(n, N) * w[(size_t)n * K + k];
    out[(size_t)blockIdx.x * ldo + k] = __float2half_rn(v);
  }
}
// activations: x (B, T, K) batch-major -> out (T, Bp, ldo) time-major, term 0 and 2 = hi, term 1 = lo (split3_slot); the
// rows of the padding clips are zero.  A (16 clips x 4 steps) tile goes through shared memory so that both sides move in long
// contiguous runs (reads: 4 steps of one clip; writes: 16 clips of one step = 16 * ldo halves).
constexpr int SPL_CB = 16, SPL_TB = 4;
__global__ void __launch_bounds__(256)
split3_f16_tm_kernel(const float* __restrict__ x, __half* __restrict__ out, int B, int Bp, int T, int K, int ldo) {
  extern __shared__ float spl_tile[];          // [SPL_CB][SPL_TB * K]
  const int b0 = blockIdx.x * SPL_CB, t0 = blockIdx.y * SPL_TB;
  const int nt = min(SPL_TB, T - t0);
  const int run = nt * K;                      // contiguous floats per clip
  for (int idx = threadIdx.x; idx < SPL_CB * run; idx += blockDim.x) {
    const int i = idx / run, j = idx % run;
    spl_tile[i * (SPL_TB * K) + j] = b0 + i < B ? __ldg(x + ((size_t)(b0 + i) * T + t0) * K + j) : 0.f;
  }
  __syncthreads();
  const int half2_per_row = ldo / 2;
  for (int idx = threadIdx.x; idx < nt * SPL_CB * half2_per_row; idx += blockDim.x) {
    const int k2 = idx % half2_per_row;
    const int i = (idx / half2_per_row) % SPL_CB, tt = idx / (half2_per_row * SPL_CB);
    __half v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = 2 * k2 + u;
      __half o = __float2half_rn(0.f);
      int term, feat;
      if (b0 + i < B && split3_slot(k, K, term, feat)) {
        const float f = spl_tile[i * (SPL_TB * K) + tt * K + feat];
        const __half hi = __float2half_rn(f);
        o = term == 1 ? __float2half_rn(f - __half2float(hi)) : hi;
      }
      v[u] = o;
    }
    if (b0 + i < Bp)
      reinterpret_cast<__half2*>(out + ((size_t)(t0 + tt) * Bp + b0 + i) * ldo)[k2] = __halves2half2(v[0], v[1]);
  }
}
static inline int pad8(int k) { return (k + 7) & ~7; }   // fp16 row pitch: multiple of 16 bytes (TMA)

}  // namespace avhot

using namespace avhot;

// The front-end's 40 MFCC are its 64 log-mel values times the DCT matrix D (frontend.cu; a3's input is
// [log-mel | MFCC]), so on the operand avhot_frontend_ex writes, x W_ih^T = lm (W[:, :64] + W[:, 64:] D)^T: layer 0 needs
// only the three log-mel k-blocks [lm_hi | lm_lo | lm_hi] (K = 192 instead of 320) against W' split the same way.
// W' is formed in double, once per handle (fp32-class like the 3-term product itself); output row p = packed row of the
// fused layer kernel.
__global__ void fold_dct_kernel(const float* __restrict__ w_ih0, const float* __restrict__ dct, __half* __restrict__ out, int H) {
  constexpr int NM = 64, NC = 40;
  const int p = blockIdx.x, j = threadIdx.x;               // packed row, log-mel column
  const int n = lstm_fused_src_row(p, H);
  const double sc = (n / H) == 2 ? 1.0 : 0.5;              // sigmoid gates halved (tc_gate_scale)
  double v = w_ih0[(size_t)n * (NM + NC) + j];
  for (int c = 0; c < NC; ++c) v += (double)w_ih0[(size_t)n * (NM + NC) + NM + c] * (double)dct[c * NM + j];
  const float f = (float)(sc * v);
  const __half hi = __float2half_rn(f);
  out[(size_t)p * 192 + j] = hi;
  out[(size_t)p * 192 + 64 + j] = hi;
  out[(size_t)p * 192 + 128 + j] = __float2half_rn(f - __half2float(hi));
}
static int lstm_build_dct_fold(avhot_lstm* h, const float* w_ih0_dev, float* dct_scratch_dev, cudaStream_t s) {
  std::vector<float> dct(40 * 64);
  int rc = avhot_frontend_tables_host(nullptr, dct.data(), nullptr);   // the table the front-end kernel was built from
  if (rc != AVHOT_OK) return rc;
  AVHOT_CUDA(cudaMemcpyAsync(dct_scratch_dev, dct.data(), sizeof(float) * dct.size(), cudaMemcpyHostToDevice, s));
  AVHOT_CUDA(cudaStreamSynchronize(s));                    // `dct` is pageable host memory
  fold_dct_kernel<<<4 * h->hidden_dim, 64, 0, s>>>(w_ih0_dev, dct_scratch_dev, reinterpret_cast<__half*>(h->w_ih_fold_p16),
                                                   h->hidden_dim);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

extern "C" int avhot_lstm_create(const avhot_lstm_desc* d, avhot_lstm_t* out, avhot_stream_t stream) {
  AVHOT_REQUIRE(d && out, AVHOT_ERR_INVALID, "lstm_create: null argument");
  AVHOT_REQUIRE(d->num_layers >= 1 && d->num_layers <= 4, AVHOT_ERR_INVALID,
                "lstm_create: num_layers=%d (1..4)", d->num_layers);
  AVHOT_REQUIRE(d->input_dim > 0 && d->output_dim > 0, AVHOT_ERR_INVALID, "lstm_create: bad dims");
  AVHOT_REQUIRE(d->hidden_dim == LH, AVHOT_ERR_UNSUPPORTED,
                "lstm_create: hidden_dim=%d; this build is specialised for hidden_dim=256 "
                "(configs/base.yaml:35)", d->hidden_dim);
  AVHOT_REQUIRE((d->input_dim & 3) == 0, AVHOT_ERR_UNSUPPORTED,
                "lstm_create: input_dim=%d must be a multiple of 4", d->input_dim);
  const int H = LH, L = d->num_layers;
  size_t total = 0;
  for (int l = 0; l < L; ++l) {
    const int in = l == 0 ? d->input_dim : H;
    AVHOT_REQUIRE(d->weight_ih[l] && d->weight_hh[l] && d->bias_ih[l] && d->bias_hh[l], AVHOT_ERR_INVALID,
                  "lstm_create: layer %d tensors missing", l);
    total += align256(sizeof(float) * 4 * H * in) + align256(sizeof(float) * 4 * H * H) +
             2 * align256(sizeof(float) * 4 * H) + align256(2 * 4 * H * H);
  }
  AVHOT_REQUIRE(d->proj_weight && d->proj_bias, AVHOT_ERR_INVALID, "lstm_create: projection missing");
  total += align256(sizeof(float) * d->output_dim * H) + align256(sizeof(float) * d->output_dim);
  total += align256(2 * 4 * H * pad8(3 * d->input_dim + 2)) + (L - 1) * align256(2 * 4 * H * H);
  // the same in the fused layer kernel's row order: W_ih (layer 0: split), W_hh, bias
  total += align256(2 * 4 * H * pad8(3 * d->input_dim + 2)) + (L - 1) * align256(2 * 4 * H * H) +
           L * (align256(2 * 4 * H * H) + align256(sizeof(float) * 4 * H));
  avhot_lstm* h = new avhot_lstm();
  h->input_dim = d->input_dim; h->hidden_dim = H; h->num_layers = L; h->output_dim = d->output_dim;
  h->arena = nullptr;
  h->w_ih_fold_p16 = nullptr;
  total += align256(2 * 4 * H * 192) + align256(sizeof(float) * 40 * 64);
  cudaError_t e = cudaMalloc(&h->arena, total);
  if (e != cudaSuccess) {
    delete h;
    set_error("lstm_create: cudaMalloc(%zu) -> %s", total, cudaGetErrorString(e));
    return AVHOT_ERR_CUDA;
  }
  cudaStream_t s = as_stream(stream);
  char* p = reinterpret_cast<char*>(h->arena);
  auto take = [&](size_t bytes) { float* r = reinterpret_cast<float*>(p); p += align256(bytes); return r; };
  int rc = AVHOT_OK;
  for (int l = 0; l < L && rc == AVHOT_OK; ++l) {
    const int in = l == 0 ? d->input_dim : H;
    h->w_ih[l] = take(sizeof(float) * 4 * H * in);
    h->w_hh_t[l] = take(sizeof(float) * 4 * H * H);
    h->bias[l] = take(sizeof(float) * 4 * H);
    h->bias_tc[l] = take(sizeof(float) * 4 * H);
    h->w_hh_r16[l] = take(2 * 4 * H * H);
    if (cudaMemcpyAsync(h->w_ih[l], d->weight_ih[l], sizeof(float) * 4 * H * in, cudaMemcpyDeviceToDevice, s) !=
        cudaSuccess) rc = AVHOT_ERR_CUDA;
    if (rc == AVHOT_OK) {                                   // [k][j][gate], once per handle
      regroup_whh_kernel<<<H, LH, 0, s>>>(d->weight_hh[l], h->w_hh_t[l]);
      if (cudaGetLastError() != cudaSuccess) rc = AVHOT_ERR_CUDA;
    }
    if (rc == AVHOT_OK) rc = add_vec(d->bias_ih[l], d->bias_hh[l], h->bias[l], 4 * H, s);
    if (rc == AVHOT_OK) rc = lstm_pack_whh_fp16(d->weight_hh[l], h->w_hh_r16[l], s);
  }
  for (int l = 0; l < L && rc == AVHOT_OK; ++l) {   // after the bias sums above (same stream)
    scale_gate_bias_kernel<<<ceil_div(4 * H, 256), 256, 0, s>>>(h->bias[l], h->bias_tc[l], 4 * H, false);
    const int ldo = l == 0 ? pad8(3 * d->input_dim + 2) : H;
    h->w_ih_f16[l] = take(2 * 4 * H * ldo);
    h->w_ih_p16[l] = take(2 * 4 * H * ldo);
    h->w_hh_p16[l] = take(2 * 4 * H * H);
    h->bias_p[l] = take(sizeof(float) * 4 * H);
    for (int perm = 0; perm < 2; ++perm) {
      __half* dst = reinterpret_cast<__half*>(perm ? h->w_ih_p16[l] : h->w_ih_f16[l]);
      if (l == 0) pack_w_split3_kernel<<<4 * H, 128, 0, s>>>(d->weight_ih[0], h->bias[0], dst, 4 * H, d->input_dim, ldo, perm);
      else pack_w_bias_kernel<<<4 * H, 128, 0, s>>>(d->weight_ih[l], h->bias[l], dst, 4 * H, H, ldo, perm);
    }
    pack_w_bias_kernel<<<4 * H, 128, 0, s>>>(d->weight_hh[l], h->bias[l], reinterpret_cast<__half*>(h->w_hh_p16[l]), 4 * H,
                                             H, H, true);
    scale_gate_bias_kernel<<<ceil_div(4 * H, 256), 256, 0, s>>>(h->bias[l], h->bias_p[l], 4 * H, true);
    if (cudaGetLastError() != cudaSuccess) rc = AVHOT_ERR_CUDA;
  }
  if (rc == AVHOT_OK && d->input_dim == 104) {
    h->w_ih_fold_p16 = take(2 * 4 * H * 192);
    rc = lstm_build_dct_fold(h, d->weight_ih[0], take(sizeof(float) * 40 * 64), s);
  }
  if (rc == AVHOT_OK) {
    h->proj_w = take(sizeof(float) * d->output_dim * H);
    h->proj_b = take(sizeof(float) * d->output_dim);
    if (cudaMemcpyAsync(h->proj_w, d->proj_weight, sizeof(float) * d->output_dim * H, cudaMemcpyDeviceToDevice, s) !=
            cudaSuccess ||
        cudaMemcpyAsync(h->proj_b, d->proj_bias, sizeof(float) * d->output_dim, cudaMemcpyDeviceToDevice, s) !=
            cudaSuccess)
      rc = AVHOT_ERR_CUDA;
  }
  if (rc != AVHOT_OK) {
    if (rc == AVHOT_ERR_CUDA) set_error("lstm_create: device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(h->arena);
    delete h;
    return rc;
  }
  *out = h;
  return AVHOT_OK;
}

extern "C" void avhot_lstm_destroy(avhot_lstm_t h) {
  if (!h) return;
  cudaFree(h->arena);
  delete h;
}

// workspace (fp32 path): xp (B*T*4H) + h_seq (B*T*H) + h_last (B*H)
// workspace: xp (Bp*T*4H fp32; none on the TC path when every layer runs fused) + h_seq (Bp*T*H fp32 / fp16) + h_last (B*H) [+ fp16 split of the input, TC path];
// Bp = B rounded up to the 128-clip tile of the tensor-core kernels
// does every layer of the stack run as the fused kernel (lstm_fused.cu) at this batch?  Then no xp buffer is needed.
static bool lstm_all_fused(avhot_lstm_t h, int B) {
  for (int l = 0; l < h->num_layers; ++l)
    if (!lstm_fused_tc_supported(l == 0 ? pad8(3 * h->input_dim + 2) : h->hidden_dim, l + 1 < h->num_layers, B)) return false;
  return true;
}
static size_t lstm_xp_bytes(avhot_lstm_t h, size_t Bp, int B, int T, int precision) {
  if (precision == AVHOT_PREC_TC && lstm_all_fused(h, B)) return 0;
  return align256(sizeof(float) * Bp * T * 4 * h->hidden_dim);
}
extern "C" size_t avhot_lstm_workspace_bytes(avhot_lstm_t h, int B, int T, int precision) {
  if (!h || B <= 0 || T <= 0) return 0;
  const size_t H = h->hidden_dim;
  const size_t Bp = ((size_t)B + 127) / 128 * 128;
  size_t n = lstm_xp_bytes(h, Bp, B, T, precision) + align256(sizeof(float) * Bp * T * H) +
             align256(sizeof(float) * (size_t)B * H);
  if (precision == AVHOT_PREC_TC) n += align256(2 * Bp * T * (size_t)pad8(3 * h->input_dim + 2));   // fp16 split of the input
  return n;
}

extern "C" int avhot_lstm_split_pitch(avhot_lstm_t h) { return h ? pad8(3 * h->input_dim + 2) : 0; }

static int lstm_forward_impl(avhot_lstm_t h, const float* seq, const void* seq_split, const int32_t* lengths, float* out,
                             int B, int T, int precision, void* workspace, size_t workspace_bytes,
                             avhot_stream_t stream, bool frontend_operand = false);

extern "C" int avhot_lstm_forward(avhot_lstm_t h, const float* seq, const int32_t* lengths, float* out, int B,
                                  int T, int precision, void* workspace, size_t workspace_bytes,
                                  avhot_stream_t stream) {
  return lstm_forward_impl(h, seq, nullptr, lengths, out, B, T, precision, workspace, workspace_bytes, stream);
}

extern "C" int avhot_lstm_forward_split(avhot_lstm_t h, const void* seq_split_f16, const int32_t* lengths, float* out,
                                        int B, int T, void* workspace, size_t workspace_bytes,
                                        avhot_stream_t stream) {
  AVHOT_REQUIRE(seq_split_f16, AVHOT_ERR_INVALID, "lstm_forward_split: operand is NULL");
  return lstm_forward_impl(h, nullptr, seq_split_f16, lengths, out, B, T, AVHOT_PREC_TC, workspace, workspace_bytes,
                           stream);
}

extern "C" int avhot_lstm_forward_frontend(avhot_lstm_t h, const void* seq_split_f16, const int32_t* lengths, float* out,
                                           int B, int T, void* workspace, size_t workspace_bytes,
                                           avhot_stream_t stream) {
  AVHOT_REQUIRE(seq_split_f16, AVHOT_ERR_INVALID, "lstm_forward_frontend: operand is NULL");
  return lstm_forward_impl(h, nullptr, seq_split_f16, lengths, out, B, T, AVHOT_PREC_TC, workspace, workspace_bytes,
                           stream, true);
}

static int lstm_forward_impl(avhot_lstm_t h, const float* seq, const void* seq_split, const int32_t* lengths, float* out,
                             int B, int T, int precision, void* workspace, size_t workspace_bytes, avhot_stream_t stream,
                             bool frontend_operand) {
  AVHOT_REQUIRE(h && (seq || seq_split) && out && B >= 0 && T >= 1, AVHOT_ERR_INVALID, "lstm_forward: bad arguments");
  AVHOT_REQUIRE(precision == AVHOT_PREC_FP32 || precision == AVHOT_PREC_TC, AVHOT_ERR_UNSUPPORTED,
                "lstm_forward: precision %d not built", precision);
  if (B == 0) return AVHOT_OK;
  const size_t need = avhot_lstm_workspace_bytes(h, B, T, precision);
  AVHOT_REQUIRE(workspace && workspace_bytes >= need, AVHOT_ERR_WORKSPACE,
                "lstm_forward: workspace %zu < %zu bytes", workspace_bytes, need);
  AVHOT_REQUIRE(((size_t)B + 127) * T <= 0x7fffffffu / 8, AVHOT_ERR_INVALID, "lstm_forward: B*T too large for one call");
  const int H = h->hidden_dim;
  cudaStream_t s = as_stream(stream);
  char* p = reinterpret_cast<char*>(workspace);
  const size_t Bp = ((size_t)B + 127) / 128 * 128;
  float* xp = reinterpret_cast<float*>(p); p += lstm_xp_bytes(h, Bp, B, T, precision);   // unused when every layer is fused
  float* h_seq = reinterpret_cast<float*>(p); p += align256(sizeof(float) * Bp * T * H);
  float* h_last = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)B * H);
  __half* x_split = reinterpret_cast<__half*>(p);   // TC path only
  for (int l = 0; l < h->num_layers; ++l) {
    const float* x = l == 0 ? seq : h_seq;
    const int in = l == 0 ? h->input_dim : H;
    const bool last = l == h->num_layers - 1;
    float* hs_out = last ? nullptr : h_seq;   // layer l overwrites its own input only after xp is formed
    float* hl_out = last ? h_last : nullptr;
    if (precision == AVHOT_PREC_TC) {
      // TC path: h_seq holds the fp16 hidden sequence (it is only ever the next layer's GEMM operand)
      int rc;
      AVHOT_REQUIRE(3 * in + 2 <= 320 || l > 0, AVHOT_ERR_UNSUPPORTED, "lstm_forward(tc): input_dim %d > 104", in);
      const int Mp = (int)Bp * T;                      // time-major rows t * Bp + b
      // xp is stored in fp16 (half the stream that bounds both kernels).  Measured against the fp64 oracle the audio
      // embedding error is unchanged (2.7e-4 with fp16 xp for both layers, 3.2e-4 with fp32 xp for layer 0), although
      // layer 0's pre-activations reach |xp| ~ 20; AVHOT_XP0_F32=1 keeps layer 0's xp in fp32.
      static const bool xp0_f32 = [] { const char* e = getenv("AVHOT_XP0_F32"); return e && atoi(e) != 0; }();
      const bool xf16 = l > 0 || !xp0_f32;
      const int hp = H;                                // fp16 pitch of h_seq
      // layer 0: A = [hi | lo | hi] x W = [W_hi | W_hi | W_lo] (K = 3 in), an fp32-class product (dropping the
      // hi x W_lo term measured 5.4e-3 on the logits, outside the stated tolerance)
      constexpr int l0_terms = 3;
      // one kernel per layer when it fits (lstm_fused.cu): input projection and recurrence in one MMA chain, no xp
      const void* x_op = l > 0 ? (const void*)h_seq : seq_split ? seq_split : (const void*)x_split;
      const int x_cols = l == 0 ? pad8(3 * in + 2) : H;
      const bool fused = lstm_fused_tc_supported(x_cols, !last, B);
      static const bool no_fold = [] { const char* e = getenv("AVHOT_LSTM_DCT_FOLD"); return e && atoi(e) == 0; }();
      if (l == 0 && frontend_operand && seq_split && h->w_ih_fold_p16 && !no_fold &&
          lstm_fused_tc_supported(192, !last, B)) {
        // the operand's MFCC columns are the DCT of its log-mel columns: folded weights, the log-mel k-blocks only
        rc = lstm_layer_fused_tc(h->w_ih_fold_p16, h->w_hh_p16[0], h->bias_p[0], seq_split, 192, x_cols,
                                 last ? nullptr : (void*)h_seq, hp, hl_out, lengths, B, T, s);
        if (rc != AVHOT_OK) return rc;
        continue;
      }
      if (l == 0 && !seq_split) {                     // (operand not prepared by avhot_frontend_ex)
        ProfScope prof("lstm_split3", s);
        const size_t smem = sizeof(float) * SPL_CB * SPL_TB * in;
        AVHOT_REQUIRE(smem <= 48 * 1024, AVHOT_ERR_UNSUPPORTED, "lstm_forward(tc): input_dim %d too large", in);
        split3_f16_tm_kernel<<<dim3((unsigned)(Bp / SPL_CB), (unsigned)ceil_div(T, SPL_TB)), 256, smem, s>>>(
            x, x_split, B, (int)Bp, T, in, x_cols);
        AVHOT_LAUNCH_CHECK();
      }
      if (fused) {
        // a middle layer reads and writes h_seq in place: step t's rows are stored after every MMA that reads them
        rc = lstm_layer_fused_tc(h->w_ih_p16[l], h->w_hh_p16[l], h->bias_p[l], x_op, x_cols, x_cols,
                                 last ? nullptr : (void*)h_seq, hp, hl_out, lengths, B, T, s);
        if (rc != AVHOT_OK) return rc;
        continue;
      }
      if (l == 0) {
        rc = gemm_f16_ws(x_op, x_cols, h->w_ih_f16[0], x_cols, h->bias_tc[0], xp, xf16, 4 * H, Mp, 4 * H, l0_terms * in, s, "lstm_input_gemm_tc");
      } else {
        rc = gemm_f16_ws(h_seq, hp, h->w_ih_f16[l], hp, h->bias_tc[l], xp, true, 4 * H, Mp, 4 * H, H, s, "lstm_input_gemm_tc");
      }
      if (rc == AVHOT_OK)
        rc = lstm_recurrence_tc(h->w_hh_r16[l], xp, xf16, last ? nullptr : (void*)h_seq, hp, hl_out, lengths, B, T, s);
      if (rc != AVHOT_OK) return rc;
      continue;
    }
    int rc = gemm_f32(x, nullptr, in, h->w_ih[l], in, h->bias[l], xp, 4 * H, B * T, 4 * H, in, ACT_NONE, s,
                      "lstm_input_gemm");
    if (rc != AVHOT_OK) return rc;
    ProfScope prof("lstm_recurrence", s);
    // Clips per CTA: a CTA streams all of W_hh (1 MB) from L2 every step whatever its tile, so few large tiles cost the
    // least L2 traffic, but the 602 dependent steps want every SM busy: the smallest multiple of 4 that fits the batch
    // into one wave of CTAs (e.g. 2560 near-tie clips of a flush: 20 clips x 128 CTAs instead of 32 x 80, measured
    // 47 -> 25 ms for the two layers), 32 for large batches.
    {   // small batches: W_hh resident in a cluster's shared memory instead of streamed every step (lstm_cs.cu)
      int rc_cs = AVHOT_OK;
      if (lstm_recurrence_cs(xp, h->w_hh_t[l], hs_out, hl_out, lengths, B, T, s, &rc_cs)) {
        if (rc_cs != AVHOT_OK) return rc_cs;
        AVHOT_LAUNCH_CHECK();
        continue;
      }
    }
    const int want = 4 * ceil_div(B, 4 * num_sms());
#define AVHOT_REC_CASE(BT_)                                                                                        \
  case BT_: {                                                                                                       \
    const size_t smem = sizeof(float) * (2 * LH * (BT_ + 4) + (size_t)lstm_ring_depth(BT_) * lstm_chunk_rows(BT_) * 4 * LH); \
    AVHOT_CUDA(cudaFuncSetAttribute(lstm_recurrence_kernel<BT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                    (int)smem));                                                                    \
    lstm_recurrence_kernel<BT_><<<ceil_div(B, BT_), LH, smem, s>>>(xp, h->w_hh_t[l], hs_out, hl_out, lengths, B, T); \
    break;                                                                                                          \
  }
    switch (want >= 32 ? 32 : want) {
      AVHOT_REC_CASE(4) AVHOT_REC_CASE(8) AVHOT_REC_CASE(12) AVHOT_REC_CASE(16) AVHOT_REC_CASE(20) AVHOT_REC_CASE(24)
      AVHOT_REC_CASE(28)
      default: AVHOT_REC_CASE(32)
    }
#undef AVHOT_REC_CASE
    AVHOT_LAUNCH_CHECK();
  }
  return gemm_f32(h_last, nullptr, H, h->proj_w, H, h->proj_b, out, h->output_dim, B, h->output_dim, H, ACT_NONE, s,
                  "lstm_projection");
}
