// Shared helpers for libavhot (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/avhot.h"

namespace avhot {

void set_error(const char* fmt, ...);
void count_launch();   // every kernel launch of the library is counted (avhot_launch_count)

// Optional per-kernel timing (avhot_prof_*): CUDA events recorded on the launching stream around the
// launches of one named kernel.  A no-op unless enabled, so the product path pays one branch.
struct ProfScope {
  cudaStream_t s;
  int slot;
  ProfScope(const char* name, cudaStream_t stream);
  ~ProfScope();
};

#define AVHOT_REQUIRE(cond, code, ...)     \
  do {                                     \
    if (!(cond)) {                         \
      ::avhot::set_error(__VA_ARGS__);     \
      return (code);                       \
    }                                      \
  } while (0)

#define AVHOT_CUDA(call)                                                                \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::avhot::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                  \
                         cudaGetErrorString(e__));                                      \
      return AVHOT_ERR_CUDA;                                                            \
    }                                                                                   \
  } while (0)

#define AVHOT_LAUNCH_CHECK()                                                            \
  do {                                                                                  \
    ::avhot::count_launch();                                                            \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      ::avhot::set_error("%s:%d launch -> %s", __FILE__, __LINE__,                     \
                         cudaGetErrorString(e__));                                      \
      return AVHOT_ERR_CUDA;                                                            \
    }                                                                                   \
  } while (0)

static inline cudaStream_t as_stream(avhot_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
int num_sms();
// core.cu
int transpose(const float* src, float* dst, int R, int C, cudaStream_t s);   // dst[c][r] = src[r][c]
int add_vec(const float* a, const float* b, float* out, int n, cudaStream_t s);
// gemm.cu: C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]); fp32 FMA on CUDA cores.
// Exactly one of a_f32 / a_u8 is non-NULL (u8 pixel k means fp32(k)/255).  K % 4 == 0.
enum { ACT_NONE = 0, ACT_RELU = 1 };
int gemm_f32(const float* a_f32, const uint8_t* a_u8, int lda, const float* w, int ldw,
             const float* bias, float* c, int ldc, int M, int N, int K, int act, cudaStream_t s,
             const char* prof_name = "gemm_f32");

// gemm_tc.cu: same contract on the tcgen05 tensor cores (fp32 operands multiplied as TF32, fp32 accumulate).
int gemm_tf32_tc(const float* a, int lda, const float* w, int ldw, const float* bias, float* c, int ldc, int M,
                 int N, int K, int act, cudaStream_t s, const char* prof_name = "gemm_tf32_tc");

// gemm_u8.cu: C[M, 256] = act(A_u8[M, K] / 255 * W^T + bias) on tcgen05 with the pixels converted to fp16 in the kernel.
// wq_f16 (256, K) / inv_scale_dev come from gemm_u8_pack_weights (scratch_dev: one uint32 of device scratch).
int gemm_u8_pack_weights(const float* w, int N, int K, void* wq_f16, float* inv_scale_dev, uint32_t* scratch_dev,
                         cudaStream_t s);
bool gemm_u8_tc_supported(const uint8_t* a, int lda, int N, int K);
int gemm_u8_tc(const uint8_t* a, int lda, const void* wq_f16, const float* bias, const float* inv_scale_dev, float* c,
               int ldc, int M, int N, int K, int act, cudaStream_t s, const char* prof_name = "gemm_u8_tc");

// gemm_cvt.cu: the same Linear on fp32 frames, converted to fp16 inside the kernel, CTA pairs sharing the fp16 weights
// (wq_f16 / inv_scale_dev of gemm_u8_pack_weights; scale_mul = 255 for frames that are pixels / 255)
bool gemm_f32cvt_tc_supported(const float* a, int lda, int N, int K);
bool gemm_u8_pair_tc_supported(const uint8_t* a, int lda, int N, int K);
int gemm_u8_pair_tc(const uint8_t* a, int lda, const void* wq_f16, const float* bias, const float* inv_scale_dev, float* c,
                    int ldc, int M, int N, int K, int act, cudaStream_t s, const char* prof_name = "gemm_u8_pair_tc");
int gemm_f32cvt_tc(const float* a, int lda, const void* wq_f16, const float* bias, const float* inv_scale_dev, float scale_mul,
                   float* c, int ldc, int M, int N, int K, int act, cudaStream_t s, const char* prof_name = "gemm_f32cvt_tc");

// lstm_tc.cu: W_hh (4H, H) fp32 -> fp16, rows regrouped per 32-unit chunk; and the tcgen05 recurrence of one layer.
int lstm_pack_whh_fp16(const float* w_hh, void* out, cudaStream_t s);
// h_seq_f16: (B, T, H) fp16 hidden sequence (input of the next layer's projection) or NULL
// xp: input projections incl. bias, time-major tile-transposed (gemm_ws.cu); h_seq_f16: (T, Bp, h_pitch) fp16 or NULL
int lstm_recurrence_tc(const void* whh_packed_fp16, const void* xp, bool xp_f16, void* h_seq_f16, int h_pitch,
                       float* h_last, const int32_t* lengths, int B, int T, cudaStream_t s);
// lstm_fused.cu: one LSTM layer, input projection + recurrence in one tcgen05 kernel (h in TMEM, c in registers).
// Row order of its packed weights / bias: row p = c*128 + q*32 + hf*16 + g*4 + i holds gate g of unit c*32 + q*8 + hf*4 + i
__host__ __device__ __forceinline__ int lstm_fused_src_row(int p, int H = 256) {
  return ((p >> 2) & 3) * H + (p >> 7) * 32 + ((p >> 5) & 3) * 8 + ((p >> 4) & 1) * 4 + (p & 3);
}
bool lstm_fused_tc_supported(int x_cols, bool hseq, int B);
int lstm_layer_fused_tc(const void* w_ih_p16, const void* w_hh_p16, const float* bias_p, const void* x_f16, int x_cols,
                        int x_pitch, void* h_seq_f16, int h_pitch, float* h_last, const int32_t* lengths, int B, int T,
                        cudaStream_t s);
// lstm_cs.cu: fp32 recurrence of one layer with W_hh resident in the shared memory of an 8-CTA cluster (small batches);
// false = the batch does not fit one wave of clusters, nothing launched.  Bit-identical to lstm_recurrence_kernel.
bool lstm_recurrence_cs(const float* xp, const float* whh_t, float* h_seq, float* h_last, const int32_t* lengths, int B,
                        int T, cudaStream_t s, int* rc);
// gemm_ws.cu: weight-stationary tcgen05 GEMM, fp16 operands, fp32 output (LSTM input projections)
int gemm_f16_ws(const void* a, int lda, const void* w, int ldw, const float* bias, void* c, bool out_f16, int ldc, int M,
                int N, int K, cudaStream_t s, const char* prof_name = "gemm_f16_ws");

// ---------------------------------------------------------------------------
// Philox-4x32-10 (oracle/philox.py is the written spec; Random123 algorithm).
// ---------------------------------------------------------------------------
struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ uint32_t philox_word(const Philox4& p, uint32_t i) {
  return i == 0 ? p.x : (i == 1 ? p.y : (i == 2 ? p.z : p.w));
}

// Dense sites - masks of >= AVHOT_DENSE_SITE_ELEMS elements per (clip, pass), i.e. the (F, H) mask of FrameEncoder's
// frame_mlp dropout - draw 16 bits per element instead of 32 (oracle/philox.py): element e uses half-word e & 1
// (0 = low) of word (e >> 1) & 3 of block e >> 3 and is kept iff half >= threshold >> 16 = floor(p * 2^16).
constexpr int AVHOT_DENSE_SITE_ELEMS = 1024;
__host__ __device__ __forceinline__ uint32_t philox_half(const Philox4& p, uint32_t i) {   // i = e & 7
  const uint32_t w = philox_word(p, i >> 1);
  return (i & 1u) ? (w >> 16) : (w & 0xFFFFu);
}

// philox_half(p, i) >= thr16 without extracting the half-word: the high half of a word compares as the word itself against
// thr16 << 16, the low half as the word shifted up (one shift for the low halves, none for the high ones, instead of a
// mask / shift per element on the integer pipe that bounds frame_tail_kernel)
__device__ __forceinline__ bool philox_half_keep(const Philox4& p, uint32_t i, uint32_t thr16_hi) {   // thr16_hi = thr16 << 16
  const uint32_t w = philox_word(p, i >> 1);
  return ((i & 1u) ? w : (w << 16)) >= thr16_hi;
}

// floor(p * 2^32) clamped to uint32: an element is kept iff word >= threshold.
static inline uint32_t drop_threshold(float p) {
  double v = (double)p * 4294967296.0;
  if (v <= 0.0) return 0u;
  if (v >= 4294967295.0) return 0xFFFFFFFFu;
  return (uint32_t)v;
}
// nn.Dropout's multiplier in fp32: 1 / (1 - p)
static inline float keep_scale(float p) { return 1.0f / (1.0f - p); }

struct DropSpec {  // what a kernel needs to draw the masks of one launch
  uint32_t k0, k1;       // seed
  uint32_t thr;          // drop threshold (0 => dropout off)
  float scale;           // 1/(1-p)
  uint64_t clip0;        // global index of clip 0 of this launch
  const uint32_t* clip_ids;   // optional (device): global index of clip b = clip_ids[b] instead of clip0 + b (the fp32
                              // re-evaluation of scattered near-tie clips, avhot_*_ids)
  uint32_t t0;           // pass index of pass 0 of this launch
  uint32_t rk0[10], rk1[10];   // the ten round keys of (k0, k1): kernel parameters live in the constant bank, so the
                               // XOR of a round reads them as an operand instead of re-deriving them (philox4x32_10_rk)
};
static inline DropSpec make_drop(float p, uint64_t seed, uint64_t clip0, uint32_t t0) {
  DropSpec d;
  d.k0 = (uint32_t)seed;
  d.k1 = (uint32_t)(seed >> 32);
  d.thr = p > 0.f ? drop_threshold(p) : 0u;
  d.scale = p > 0.f ? keep_scale(p) : 1.0f;
  d.clip0 = clip0;
  d.clip_ids = nullptr;
  d.t0 = t0;
  for (int r = 0; r < 10; ++r) {
    d.rk0[r] = d.k0 + (uint32_t)r * 0x9E3779B9u;
    d.rk1[r] = d.k1 + (uint32_t)r * 0xBB67AE85u;
  }
  return d;
}
// Philox counter word of clip b of the launch (low 32 bits of its global index)
__device__ __forceinline__ uint32_t ds_clip(const DropSpec& ds, int b) {
  return ds.clip_ids ? __ldg(ds.clip_ids + b) : (uint32_t)(ds.clip0 + (uint64_t)b);
}
// philox4x32_10(c0, c1, c2, c3, ds.k0, ds.k1) with the key schedule taken from the DropSpec (a kernel parameter): saves
// the 18 key additions per block that the compiler otherwise re-issues for every block (frame_tail_kernel draws 34 M
// blocks per step).  Bit-identical results.
__device__ __forceinline__ Philox4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const DropSpec& ds) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ ds.rk0[r], n2 = hi0 ^ c3 ^ ds.rk1[r];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace avhot
