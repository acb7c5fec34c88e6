// a2: log-mel + MFCC front-end.  Written spec: oracle/frontend.py (the reference has no
// implementation of this stage).
//
//   wave (B, N) fp32 -> feats (B, 1 + N/160, 104) fp32 = [64 log-mel | 40 MFCC]
//   and / or the tensor-core LSTM's first GEMM operand (fp16 hi / lo split, time-major; avhot_frontend_ex)
//
// One persistent kernel (frontend_tc_kernel, below): framing, pre-emphasis and windowing from shared memory, a
// shared-memory radix-16 x 16 STFT with a warp-shuffle real-FFT split / power spectrum on the CUDA cores, the mel
// filterbank projection and the MFCC DCT as tcgen05 / TMEM tiles, log and scaling in the epilogues, TMA stores.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tc.cuh"

namespace avhot {

constexpr int N_FFT = 512, HOP = 160, N_BINS = 257, N_MELS = 64, N_MFCC = 40, N_FEATS = 104;

struct FrontendTables {            // device layout
  float window[N_FFT];
  float2 tw[N_FFT];                // tw[k] = exp(-2 pi i k / 512)
  // tcgen05 operands, fp16 bit patterns already in the K-major 128-byte-swizzled shared-memory layout
  // ([k-block of 64][row][16-byte chunk ^ (row & 7)]), copied to shared memory as they are:
  uint16_t b3lo[32 * 64];            //   mel weights x 64 (= 0.25 for the 4 |X|^2 of the kernel x 256 to keep small taps
  uint16_t b3hi[4 * 32 * 64];        //   normal in fp16), banded: filters 0-31 over bins 0-63 (they reach no further),
                                     //   filters 32-63 over bins 0-255; row = filter
  uint16_t b4[2 * 48 * 64];          //   DCT matrix x 64, hi then lo = fp16(D - hi): row = coefficient (40, padded to
                                     //   48), K = 64 log-mel bands
  float2 t1[256];                    // t1[c * 16 + b] = exp(-2 pi i b c / 256): twiddles between the two radix-16 passes
};

static FrontendTables g_host_tables;
static std::vector<float> g_host_mel_dense, g_host_dct;
static std::once_flag g_tables_once;
static bool g_tables_ok = false;

static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static void build_tables() {
  FrontendTables& t = g_host_tables;
  const double PI = 3.14159265358979323846;
  for (int n = 0; n < N_FFT; ++n) {
    t.window[n] = (float)(0.5 - 0.5 * cos(2.0 * PI * n / N_FFT));
    t.tw[n] = make_float2((float)cos(2.0 * PI * n / N_FFT), (float)(-sin(2.0 * PI * n / N_FFT)));
  }
  // Slaney mel filterbank, Slaney area normalisation, 0..8000 Hz (oracle/frontend.py mel_filterbank)
  const double sr = 16000.0, f_lo = 0.0, f_hi = 8000.0;
  std::vector<double> hz(N_MELS + 2);
  const double m_lo = hz_to_mel(f_lo), m_hi = hz_to_mel(f_hi);
  for (int i = 0; i < N_MELS + 2; ++i) hz[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (N_MELS + 1));
  g_host_mel_dense.assign(N_MELS * N_BINS, 0.f);
  for (int m = 0; m < N_MELS; ++m) {
    const double enorm = 2.0 / (hz[m + 2] - hz[m]);
    for (int k = 0; k < N_BINS; ++k) {
      const double f = (sr / 2.0) * k / (N_BINS - 1);
      const double lower = (f - hz[m]) / (hz[m + 1] - hz[m]);
      const double upper = (hz[m + 2] - f) / (hz[m + 2] - hz[m + 1]);
      const double w = lower < upper ? lower : upper;
      if (w > 0.0) g_host_mel_dense[m * N_BINS + k] = (float)(w * enorm);
    }
  }
  // The Nyquist bin carries no weight (the last filter ends there), so the kernel keeps bins 0..255 only.
  g_tables_ok = true;
  for (int m = 0; m < N_MELS; ++m) if (g_host_mel_dense[m * N_BINS + (N_BINS - 1)] != 0.f) g_tables_ok = false;
  g_host_dct.assign(N_MFCC * N_MELS, 0.f);
  for (int c = 0; c < N_MFCC; ++c)
    for (int j = 0; j < N_MELS; ++j) {
      double v = cos(PI * (j + 0.5) * c / N_MELS) * sqrt(2.0 / N_MELS);
      if (c == 0) v *= sqrt(0.5);
      g_host_dct[c * N_MELS + j] = (float)v;
    }
  auto f16_bits = [](float v) {
    const __half h = __float2half_rn(v);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
  };
  for (int m = 0; m < N_MELS; ++m)
    for (int k = 0; k < 256; ++k) {
      const int kb = k >> 6, chunk = (k & 63) >> 3, e = k & 7, mr = m & 31;
      const uint16_t wv = f16_bits(64.f * g_host_mel_dense[m * N_BINS + k]);
      const size_t pos = (size_t)mr * 64 + ((chunk ^ (mr & 7)) << 3) + e;
      if (m >= 32) t.b3hi[(size_t)kb * (32 * 64) + pos] = wv;
      else if (kb == 0) t.b3lo[pos] = wv;
      else if (g_host_mel_dense[m * N_BINS + k] != 0.f) g_tables_ok = false;   // the banded layout assumes it
    }
  for (int c = 0; c < 16; ++c)
    for (int bq = 0; bq < 16; ++bq) {
      const double ang = -2.0 * PI * (double)(bq * c) / 256.0;
      t.t1[c * 16 + bq] = make_float2((float)cos(ang), (float)sin(ang));
    }
  for (int c = 0; c < 48; ++c)
    for (int j = 0; j < N_MELS; ++j) {
      const float d = c < N_MFCC ? 64.f * g_host_dct[c * N_MELS + j] : 0.f;
      const uint16_t hi = f16_bits(d);
      __half hh;
      memcpy(&hh, &hi, 2);
      const size_t pos = (size_t)c * 64 + (((j >> 3) ^ (c & 7)) << 3) + (j & 7);
      t.b4[pos] = hi;
      t.b4[48 * 64 + pos] = f16_bits(d - __half2float(hh));
    }
}

constexpr int MAX_DEVICES = 64;
static FrontendTables* g_dev_tables[MAX_DEVICES] = {};
static std::mutex g_dev_mutex;

static int get_tables(const FrontendTables** out) {
  std::call_once(g_tables_once, build_tables);
  AVHOT_REQUIRE(g_tables_ok, AVHOT_ERR_UNSUPPORTED, "frontend: mel filterbank does not fit the tensor-core layout");
  int dev = 0;
  AVHOT_CUDA(cudaGetDevice(&dev));
  AVHOT_REQUIRE(dev >= 0 && dev < MAX_DEVICES, AVHOT_ERR_INVALID, "frontend: device index %d", dev);
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  if (!g_dev_tables[dev]) {
    FrontendTables* d = nullptr;
    AVHOT_CUDA(cudaMalloc(&d, sizeof(FrontendTables)));
    AVHOT_CUDA(cudaMemcpy(d, &g_host_tables, sizeof(FrontendTables), cudaMemcpyHostToDevice));
    g_dev_tables[dev] = d;
  }
  *out = g_dev_tables[dev];
  return AVHOT_OK;
}

// ---- complex helpers -------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// DFT-4, natural order in and out (forward transform, e^{-i...})
__device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = mul_mi(csub(v1, v3));
  v0 = cadd(a0, a2);
  v1 = cadd(a1, a3);
  v2 = csub(a0, a2);
  v3 = csub(a1, a3);
}
// =====================================================================================================================
// frontend_tc_kernel: the front-end with the two dense contractions on the tcgen05 tensor cores.
//
// Persistent CTA (one per SM, grid-stride over clips), 20 warps:
//   warps 0-15   FFT producers, TWO frames per warp (half-warp = frame, lane b of 16).  A group = 32 consecutive frames
//                of a clip.  Per frame: pre-emphasis + Hann window while loading the packed samples (lane b takes the
//                complex points 16 a + b), FFT-256 as 16 x 16: a radix-16 pass in registers, the W_256^{bc} twiddles,
//                ONE exchange through shared memory (16 x 16 tile, rows padded to 17: conflict-free both ways), a second radix-16
//                pass - lane c then holds Z[c + 16 d], d = 0..15 -, real-FFT split with the mirrored bin Z[256 - k]
//                fetched from lane 16 - c by warp shuffle, |X|^2.  The frame's power spectrum is scaled by a power of
//                two that puts its largest bin at 2^13..2^14, rounded to fp16 and written as ROW (frame) of the mel
//                GEMM's A operand, K-major with the 128-byte swizzle (a3: 128 frames x 256 bins); the inverse scale goes
//                to unscale[row].  The raw samples of the next group are copied global -> shared with cp.async as
//                soon as every warp has read its inputs (one buffer, two named barriers a group); groups that touch
//                the reflect padding, or unaligned waveforms, are staged with plain loads and read through an
//                index-reflecting slow path.
//   warps 16-19  tensor-core side, thread = TMEM lane = frame row of a 128-frame tile.  Warp 16 also issues the MMAs
//                (one elected lane): D3 (128 x 64, TMEM) = a3 x b3^T, the mel filterbank as two banded N = 32 GEMMs
//                (filters 0-31 only reach bins 0-63: 4 + 16 tcgen05.mma M128 N32 K16), then D4 (128 x 48) =
//                [lm_hi | lm_lo | lm_hi] x [D_hi | D_hi | D_lo]^T, the DCT-II as an error-compensated fp16 product (12
//                MMAs M128 N48 K16) over the log-mel tile the epilogue wrote.  Epilogue: tcgen05.ld the mel energies,
//                undo the scale, floor, 10 log10; the log-mel row goes out as fp16 hi / lo 64-column tiles x[0], x[1]
//                in the same swizzled K-major layout - at once the DCT's A operand and, for the tensor-core LSTM, three
//                64-column k-blocks of its first GEMM operand (hi | lo | hi, lstm.cu split3_slot): TMA stores them
//                straight from x[] into the time-major (t, clip) rows (3-D tensor map; rows past the clip's last frame
//                are clipped by the map).  After the DCT the 40 MFCC go out the same way ([mf_hi | mf_lo | mf_hi | 0]
//                = two more k-blocks).  The fp32 feature tensor (avhot_frontend) is written row by row.
// mbarriers: a3_full (16 FFT warps -> MMA), a3_free / d3_full (MMA commit -> FFT warps / epilogue), x_full (4 epilogue
// warps -> MMA), d4_full (MMA commit -> epilogue).  Parities follow the CTA's tile counter.
// Precision: power bins and mel weights carry 11 significant bits each (fp16, round to nearest: 2^-11 relative), so a
// log-mel value is within 0.0042 dB of the fp64 spec in the worst case (measured: 0.0035 dB at most, 0.0007 dB rms); the
// per-frame power-of-two scale keeps every bin within 84 dB of the frame's peak normal in fp16.
// =====================================================================================================================
constexpr int FT_FFT_WARPS = 16, FT_EPI_WARP0 = 16, FT_THREADS = 20 * 32, FT_TILE = 128, FT_GROUP = 32;
constexpr int FT_FFT_THREADS = FT_FFT_WARPS * 32;
constexpr int SEG32 = (FT_GROUP - 1) * HOP + N_FFT;   // 5472 samples a group touches
constexpr int SEG_PRE = 4;                            // samples staged in front of them (pre-emphasis reads one back)
constexpr int SEG_BUF = 5504;                         // SEG_PRE + SEG32, rounded up
constexpr int ZROW = 17;                              // exchange tile row pitch: element (c, b) at c * 17 + b - lanes b write
                                                      // consecutive slots, lanes c read slots 17 apart (34 words: distinct
                                                      // bank pairs), and every address is a lane base plus a constant

struct FtSmem {
  __half a3[4][FT_TILE * 64];        // power spectrum tile, A operand of the mel GEMM
  __half b3lo[32 * 64];              // mel weights of filters 0-31 over bins 0-63
  __half b3hi[4][32 * 64];           // mel weights of filters 32-63 over bins 0-255
  __half x[2][FT_TILE * 64];         // log-mel hi / lo (A operand of the DCT, TMA-store source), later the MFCC block
  __half b4[2][48 * 64];             // DCT matrix hi / lo, B operand
  float2 z[FT_FFT_WARPS][2][16 * ZROW];   // FFT exchange, one 16 x 16 tile per frame, rows padded to 17
  float seg[SEG_BUF];                // raw samples of a group: seg[SEG_PRE + i] = y[group's first sample + i]
  float win[N_FFT];
  float2 tw[N_FFT / 2];              // W_512^k
  float2 t1[256];                    // t1[c * 16 + b] = W_256^{b c}
  float unscale[FT_TILE];
  uint64_t a3_full, a3_free, d3_full, x_full, d4_full;
  uint32_t tmem_base;
};
static_assert(sizeof(FtSmem) + 1024 <= 227 * 1024, "one CTA per SM");

__device__ __forceinline__ void bar_named(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void fe_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}

// DFT-16, natural order in and out: a = a1 + 4 a2, c = c2 + 4 c1, W16^{ac} = W16^{a1 c2} W4^{a1 c1} W4^{a2 c2}
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
  for (int a1 = 0; a1 < 4; ++a1) dft4(v[a1], v[a1 + 4], v[a1 + 8], v[a1 + 12]);   // v[a1 + 4 c2] = t[a1][c2]
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r = 0.70710678118654752f;
  auto mulw = [](float2 a, float wr, float wi) { return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr); };
  v[1 + 4] = mulw(v[1 + 4], c1, -s1);                                   // W16^1
  v[1 + 8] = make_float2(r * (v[1 + 8].x + v[1 + 8].y), r * (v[1 + 8].y - v[1 + 8].x));      // W16^2 = (1 - i) / sqrt 2
  v[1 + 12] = mulw(v[1 + 12], s1, -c1);                                 // W16^3
  v[2 + 4] = make_float2(r * (v[2 + 4].x + v[2 + 4].y), r * (v[2 + 4].y - v[2 + 4].x));      // W16^2
  v[2 + 8] = mul_mi(v[2 + 8]);                                          // W16^4 = -i
  v[2 + 12] = make_float2(r * (v[2 + 12].y - v[2 + 12].x), -r * (v[2 + 12].x + v[2 + 12].y));   // W16^6 = (-1 - i) / sqrt 2
  v[3 + 4] = mulw(v[3 + 4], s1, -c1);                                   // W16^3
  v[3 + 8] = make_float2(r * (v[3 + 8].y - v[3 + 8].x), -r * (v[3 + 8].x + v[3 + 8].y));       // W16^6
  v[3 + 12] = mulw(v[3 + 12], -c1, s1);                                 // W16^9
#pragma unroll
  for (int c2 = 0; c2 < 4; ++c2) dft4(v[4 * c2], v[4 * c2 + 1], v[4 * c2 + 2], v[4 * c2 + 3]);   // v[c1 + 4 c2] = Y[c2 + 4 c1]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i + 1; j < 4; ++j) { const float2 t = v[i + 4 * j]; v[i + 4 * j] = v[j + 4 * i]; v[j + 4 * i] = t; }
}

// fp16 hi / lo split of 8 floats -> two 16-byte chunks
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 hh = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * e] - hf.x, x[2 * e + 1] - hf.y);
    h[e] = *reinterpret_cast<const uint32_t*>(&hh);
    l[e] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(FT_THREADS, 1)
frontend_tc_kernel(const float* __restrict__ wave, float* __restrict__ feats, const __grid_constant__ CUtensorMap tmX,
                   int has_xs, int B, int n_samples, int n_frames, const FrontendTables* __restrict__ tab) {
  using namespace tc;
  extern __shared__ unsigned char ft_smem_raw[];
  FtSmem& sm = *reinterpret_cast<FtSmem*>(ft_smem_raw + ((1024u - (smem_u32(ft_smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int n_tiles = (n_frames + FT_TILE - 1) / FT_TILE;

  // ---- one-time setup: tables and tensor-core B operands -> shared memory, barriers, TMEM ---------------------------
  for (int i = tid; i < N_FFT; i += FT_THREADS) sm.win[i] = tab->window[i];
  for (int i = tid; i < N_FFT / 2; i += FT_THREADS) sm.tw[i] = tab->tw[i];
  for (int i = tid; i < 256; i += FT_THREADS) sm.t1[i] = tab->t1[i];
  {
    const uint4* s3 = reinterpret_cast<const uint4*>(tab->b3lo);     // b3lo and b3hi are adjacent in both structs
    uint4* d3 = reinterpret_cast<uint4*>(&sm.b3lo[0]);
    for (int i = tid; i < (int)((sizeof(sm.b3lo) + sizeof(sm.b3hi)) / 16); i += FT_THREADS) d3[i] = __ldg(s3 + i);
    const uint4* s4 = reinterpret_cast<const uint4*>(tab->b4);
    uint4* d4 = reinterpret_cast<uint4*>(&sm.b4[0][0]);
    for (int i = tid; i < (int)(sizeof(sm.b4) / 16); i += FT_THREADS) d4[i] = __ldg(s4 + i);
    fence_proxy_async_smem();                       // the tensor cores read them through the async proxy
  }
  if (tid == 0) {
    mbar_init(&sm.a3_full, FT_FFT_WARPS);
    mbar_init(&sm.a3_free, 1);
    mbar_init(&sm.d3_full, 1);
    mbar_init(&sm.x_full, 4);
    mbar_init(&sm.d4_full, 1);
    fence_barrier_init();
    if (has_xs) tma_prefetch_desc(&tmX);
  }
  if (warp == FT_EPI_WARP0) {
    tmem_alloc(&sm.tmem_base, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp < FT_FFT_WARPS) {
    // =============================== FFT producers ===================================================================
    const int h = lane >> 4, b = lane & 15;
    const int fr = 2 * warp + h;                     // this half-warp's frame of the group
    const int sw = fr & 7;                           // row & 7 (groups start at multiples of 32 rows)
    float2* const zf = sm.z[warp][h];
    const bool can_vec = (n_samples & 3) == 0 && (reinterpret_cast<uintptr_t>(wave) & 15) == 0;
    // samples of the group starting at frame0 -> seg (raw; zeros outside the clip): seg[SEG_PRE + i] = y[q0 - 256 + i]
    auto stage = [&](int clip, int frame0) {
      const float* y = wave + (size_t)clip * n_samples;
      const int n0 = frame0 * HOP - N_FFT / 2 - SEG_PRE;              // sample index of seg[0]; a multiple of 4
      for (int j = tid; j < (SEG_PRE + SEG32) / 4; j += FT_FFT_THREADS) {
        const int n = n0 + 4 * j;
        if (can_vec && n >= 0 && n + 4 <= n_samples) {
          fe_cp_async16(&sm.seg[4 * j], y + n);
        } else {
          float4 q;
          q.x = n >= 0 && n < n_samples ? __ldg(y + n) : 0.f;
          q.y = n + 1 >= 0 && n + 1 < n_samples ? __ldg(y + n + 1) : 0.f;
          q.z = n + 2 >= 0 && n + 2 < n_samples ? __ldg(y + n + 2) : 0.f;
          q.w = n + 3 >= 0 && n + 3 < n_samples ? __ldg(y + n + 3) : 0.f;
          *reinterpret_cast<float4*>(&sm.seg[4 * j]) = q;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t tcount = 0;
    stage(blockIdx.x, 0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    bar_named(2, FT_FFT_THREADS);
    // per-lane byte offsets of the power-spectrum stores inside a row of a3 (k = b + 16 d, its mirror 256 - k)
    const uint32_t base_k = ((uint32_t)((b >> 3) ^ sw) << 4) | ((uint32_t)(b & 7) << 1);
    const int bm = (16 - b) & 15;                    // lane whose registers hold Z[256 - k]
    const uint32_t base_m = ((uint32_t)((bm >> 3) ^ sw) << 4) | ((uint32_t)(bm & 7) << 1);
    const int z0 = b == 0 ? 1 : 0;
    const int src_lane = (lane & 16) | bm;
    const float2* const w2 = reinterpret_cast<const float2*>(sm.win);
    for (int clip = blockIdx.x; clip < B; clip += gridDim.x) {
      for (int tile = 0; tile < n_tiles; ++tile, ++tcount) {
        const int t_tile = tile * FT_TILE;
        const int n_groups = min(FT_TILE / FT_GROUP, (n_frames - t_tile + FT_GROUP - 1) / FT_GROUP);
        for (int g = 0; g < n_groups; ++g) {
          const int frame0 = t_tile + g * FT_GROUP;
          int nclip = clip, nframe0 = frame0 + FT_GROUP;           // where the next group starts
          if (nframe0 >= n_frames) { nclip = clip + gridDim.x; nframe0 = 0; }
          const bool has_next = nclip < B;
          const int row = g * FT_GROUP + fr;
          const int t = frame0 + fr;
          const bool valid = t < n_frames;
          const bool warp_live = frame0 + 2 * warp < n_frames;      // warp-uniform: at least the first frame exists
          float2 v[16];
          if (warp_live) {
            // ---- load: pre-emphasis + window, complex point m = 16 a + b = samples 2 m, 2 m + 1 of the frame
            const int q0 = t * HOP;                                 // centre of the frame
            const bool edge = valid && (q0 < N_FFT / 2 || q0 + N_FFT / 2 > n_samples);
            if (!__any_sync(0xffffffffu, edge)) {
              const float* fs = sm.seg + SEG_PRE + fr * HOP;
#pragma unroll
              for (int a = 0; a < 16; ++a) {
                const int m = 16 * a + b;
                const float2 s = *reinterpret_cast<const float2*>(fs + 2 * m);
                const float pm = fs[2 * m - 1];
                const float2 w = w2[m];
                v[a] = make_float2(fmaf(-0.97f, pm, s.x) * w.x, fmaf(-0.97f, s.x, s.y) * w.y);
              }
            } else {
              // frames that touch the reflect padding (numpy "reflect": no edge repeat): q[i] = p[reflect(i)],
              // p[n] = y[n] - 0.97 y[n - 1], y[-1] = 0; y[n] = seg[SEG_PRE + n - (first sample of the group)]
              const int gbase = frame0 * HOP - N_FFT / 2;
#pragma unroll
              for (int a = 0; a < 16; ++a) {
                const int m = 16 * a + b;
                float pv[2];
#pragma unroll
                for (int e2 = 0; e2 < 2; ++e2) {
                  int n = q0 - N_FFT / 2 + 2 * m + e2;
                  if (n < 0) n = -n;
                  if (n >= n_samples) n = 2 * (n_samples - 1) - n;
                  const int idx = min(max(n - gbase, 1 - SEG_PRE), SEG32 - 1);
                  const float yn = sm.seg[SEG_PRE + idx];
                  const float yp = n >= 1 ? sm.seg[SEG_PRE + idx - 1] : 0.f;
                  pv[e2] = fmaf(-0.97f, yp, yn);
                }
                const float2 w = w2[m];
                v[a] = make_float2(pv[0] * w.x, pv[1] * w.y);
              }
            }
          }
          bar_named(2, FT_FFT_THREADS);                             // every warp has read its samples: seg is free
          if (has_next) stage(nclip, nframe0);
          if (warp_live) {
            // ---- FFT-256 = 16 x 16
            dft16(v);                                               // over a: Y_b[c]
#pragma unroll
            for (int c = 1; c < 16; ++c) v[c] = cmul(v[c], sm.t1[c * 16 + b]);
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 16; ++c) zf[c * ZROW + b] = v[c];
            __syncwarp();
#pragma unroll
            for (int bb = 0; bb < 16; ++bb) v[bb] = zf[b * ZROW + bb];       // lane c = b reads (c, bb)
            dft16(v);                                               // over b: Z[c + 16 d] = v[d]
            // ---- real-FFT split + power: bins k = c + 16 d (d < 8) and 256 - k from the pair (Z_k, Z_{256-k});
            //      Z_{256-k} is register 15 - d of lane 16 - c (register 16 - d of lane 0 itself for c = 0)
            float pa[8], pb[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) {
              const float2 snd = v[15 - d];
              const float px = __shfl_sync(0xffffffffu, snd.x, src_lane), py = __shfl_sync(0xffffffffu, snd.y, src_lane);
              const float2 own = d == 0 ? v[0] : v[16 - d];
              const float2 zc = z0 ? make_float2(own.x, -own.y) : make_float2(px, -py);     // conj(Z[256 - k])
              const float2 zk = v[d];
              const float2 xe = cadd(zk, zc);
              const float2 T = cmul(sm.tw[b + 16 * d], mul_mi(csub(zk, zc)));
              const float2 Xa = cadd(xe, T), Xb = csub(xe, T);
              pa[d] = Xa.x * Xa.x + Xa.y * Xa.y;                     // 4 |X_k|^2 (the factor lives in the mel weights)
              pb[d] = Xb.x * Xb.x + Xb.y * Xb.y;                     // 4 |X_{256-k}|^2
            }
            const float p128 = 4.f * (v[8].x * v[8].x + v[8].y * v[8].y);   // lane 0: Z_128 is its own mirror
            // the frame's largest bin (bin 256 - lane 0's pb[0] - carries no filter weight and is left out)
            float mx = z0 ? p128 : pb[0];
#pragma unroll
            for (int d = 0; d < 8; ++d) mx = fmaxf(mx, fmaxf(pa[d], d ? pb[d] : mx));
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            // scale = 2^(13 - e), e = exponent of the maximum: the largest bin lands in [2^13, 2^14)
            const int e = max((int)(__float_as_uint(mx) >> 23) - 127, -100);
            const float sc = mx > 0.f ? __uint_as_float((uint32_t)(127 + 13 - e) << 23) : 1.f;
            if (g == 0) mbar_wait(&sm.a3_free, (tcount & 1u) ^ 1u);   // the previous tile's mel MMAs have read a3
            if (valid) {
              unsigned char* arow = reinterpret_cast<unsigned char*>(&sm.a3[0][0]) + row * 128;
#pragma unroll
              for (int d = 0; d < 8; ++d) {
                // bin k = b + 16 d: k-block d >> 2, chunk (b >> 3) + 2 (d & 3)
                *reinterpret_cast<__half*>(arow + (d >> 2) * (FT_TILE * 128) + (base_k ^ (uint32_t)((d & 3) << 5))) =
                    __float2half_rn(pa[d] * sc);
                // bin 256 - k = bm + 16 dd, dd = 15 - d (16 - d for lane 0, whose d = 0 mirror is the unused bin 256)
                const int dd = 15 - d + z0;
                if (dd < 16)
                  *reinterpret_cast<__half*>(arow + (dd >> 2) * (FT_TILE * 128) + (base_m ^ (uint32_t)((dd & 3) << 5))) =
                      __float2half_rn(pb[d] * sc);
              }
              if (z0) {
                *reinterpret_cast<__half*>(arow + 2 * (FT_TILE * 128) + ((uint32_t)sw << 4)) = __float2half_rn(p128 * sc);
                // mel = acc / (256 sc): the operand is 4 |X|^2 sc and the weights carry a factor 64
                sm.unscale[row] = mx > 0.f ? __uint_as_float((uint32_t)(127 + e - 13 - 8) << 23) : 1.f;
              }
            }
          } else if (g == 0) {
            mbar_wait(&sm.a3_free, (tcount & 1u) ^ 1u);
          }
          if (g == n_groups - 1) {                       // this warp's rows of the tile are in shared memory
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.a3_full);
          }
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          bar_named(2, FT_FFT_THREADS);                  // the next group's samples are in seg
        }
      }
    }
  } else {
    // =============================== tensor-core side: thread = frame row of the tile ==============================
    const int q = warp & 3, row = q * 32 + lane, et = tid - FT_EPI_WARP0 * 32;
    const bool issuer = warp == FT_EPI_WARP0;            // this warp also issues the MMAs (whole warp, one elected lane)
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    const int sw = row & 7;
    constexpr uint32_t idesc_mel = umma_idesc(FMT_F16, FT_TILE, 32);
    constexpr uint32_t idesc_dct = umma_idesc(FMT_F16, FT_TILE, 48);
    unsigned char* const x0row = reinterpret_cast<unsigned char*>(&sm.x[0][0]) + row * 128;
    unsigned char* const x1row = reinterpret_cast<unsigned char*>(&sm.x[1][0]) + row * 128;
    uint32_t tcount = 0;
    for (int clip = blockIdx.x; clip < B; clip += gridDim.x)
      for (int tile = 0; tile < n_tiles; ++tile, ++tcount) {
        const int t0 = tile * FT_TILE;
        const bool valid = t0 + row < n_frames;
        float* const frow = feats ? feats + ((size_t)clip * n_frames + t0 + row) * N_FEATS : nullptr;
        if (issuer) {
          mbar_wait_sleep(&sm.a3_full, tcount & 1u);
          tc_fence_after();
          {   // filters 0-31: bins 0-63 only
            const uint64_t ad = umma_desc_sw128(smem_u32(sm.a3[0])), bd = umma_desc_sw128(smem_u32(sm.b3lo));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_elect(tmem, ad + 2 * k, bd + 2 * k, idesc_mel, k != 0);
          }
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {   // filters 32-63
            const uint64_t ad = umma_desc_sw128(smem_u32(sm.a3[kb])), bd = umma_desc_sw128(smem_u32(sm.b3hi[kb]));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_elect(tmem + 32, ad + 2 * k, bd + 2 * k, idesc_mel, (kb | k) != 0);
          }
          umma_commit_elect(&sm.a3_free);
          umma_commit_elect(&sm.d3_full);
        }
        mbar_wait_sleep(&sm.d3_full, tcount & 1u);
        tc_fence_after();
        const float us = sm.unscale[row];
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[32];
          tmem_ld32(tl + (uint32_t)(hh * 32), r);
          tmem_ld_wait();
          float lm[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) lm[j] = 3.01029995663981195f * log2f(fmaxf(__uint_as_float(r[j]) * us, 1e-10f));
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 hi, lo;
            split8(lm + 8 * c, hi, lo);
            const int off = ((hh * 4 + c) ^ sw) << 4;
            *reinterpret_cast<uint4*>(x0row + off) = hi;
            *reinterpret_cast<uint4*>(x1row + off) = lo;
          }
          if (frow && valid) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<float4*>(frow + hh * 32 + 4 * c) = make_float4(lm[4 * c], lm[4 * c + 1], lm[4 * c + 2], lm[4 * c + 3]);
          }
        }
        tc_fence_before();                                 // D3 has been read: the next tile's mel MMAs may overwrite it
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.x_full);
        if (issuer) {
          mbar_wait_sleep(&sm.x_full, tcount & 1u);
          tc_fence_after();
#pragma unroll
          for (int term = 0; term < 3; ++term) {           // hi x D_hi + lo x D_hi + hi x D_lo
            const uint64_t ad = umma_desc_sw128(smem_u32(sm.x[term == 1 ? 1 : 0]));
            const uint64_t bd = umma_desc_sw128(smem_u32(sm.b4[term == 2 ? 1 : 0]));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_elect(tmem + 64, ad + 2 * k, bd + 2 * k, idesc_dct, (term | k) != 0);
          }
          umma_commit_elect(&sm.d4_full);
        }
        if (has_xs) {                                      // log-mel k-blocks of the LSTM operand: hi | lo | hi
          bar_named(1, 128);
          if (et == 0) {
            tma_store_3d(&tmX, sm.x[0], 0, clip, t0);
            tma_store_3d(&tmX, sm.x[1], 64, clip, t0);
            tma_store_3d(&tmX, sm.x[0], 128, clip, t0);
            tma_store_commit();
          }
        }
        mbar_wait_sleep(&sm.d4_full, tcount & 1u);
        tc_fence_after();
        float mf[40];
        {
          uint32_t a[32], bq[16];
          tmem_ld32(tl + 64u, a);
          tmem_ld16(tl + 96u, bq);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) mf[j] = __uint_as_float(a[j]) * 0.015625f;
#pragma unroll
          for (int j = 0; j < 8; ++j) mf[32 + j] = __uint_as_float(bq[j]) * 0.015625f;
        }
        tc_fence_before();
        if (frow && valid) {
#pragma unroll
          for (int c = 0; c < 10; ++c)
            *reinterpret_cast<float4*>(frow + N_MELS + 4 * c) = make_float4(mf[4 * c], mf[4 * c + 1], mf[4 * c + 2], mf[4 * c + 3]);
        }
        if (has_xs) {
          // x[] is free once the DCT's MMAs are done (d4_full) and the three stores above have read it
          if (et == 0) tma_store_wait_read0();
          bar_named(1, 128);
          // MFCC block of the operand: [mf_hi(40) | mf_lo(40) | mf_hi(40) | 0(8)] = 16 chunks over x[0], x[1]
          uint4 hi[5], lo[5];
#pragma unroll
          for (int c = 0; c < 5; ++c) split8(mf + 8 * c, hi[c], lo[c]);
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const uint4 val = c < 5 ? hi[c] : (c < 10 ? lo[c - 5] : (c < 15 ? hi[c - 10] : make_uint4(0u, 0u, 0u, 0u)));
            *reinterpret_cast<uint4*>((c < 8 ? x0row : x1row) + (((c & 7) ^ sw) << 4)) = val;
          }
          fence_proxy_async_smem();
          bar_named(1, 128);
          if (et == 0) {
            tma_store_3d(&tmX, sm.x[0], 192, clip, t0);
            tma_store_3d(&tmX, sm.x[1], 256, clip, t0);
            tma_store_commit();
            tma_store_wait_read0();
          }
          bar_named(1, 128);
        }
      }
    if (et == 0) tma_store_wait0();                        // all bulk stores have landed before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FT_EPI_WARP0) tmem_dealloc(tmem, 128);
}

}  // namespace avhot

using namespace avhot;

static int frontend_launch(const float* wave, float* feats, void* xs, int Bp, int ldx, int B, int n_samples,
                           avhot_stream_t stream) {
  AVHOT_REQUIRE(wave && (feats || xs) && B >= 0, AVHOT_ERR_INVALID, "frontend: null pointer or B < 0");
  AVHOT_REQUIRE(n_samples > N_FFT / 2, AVHOT_ERR_INVALID,
                "frontend: n_samples=%d must exceed the reflect pad (256)", n_samples);
  AVHOT_REQUIRE(B <= 65535, AVHOT_ERR_INVALID, "frontend: B > 65535 per call");
  AVHOT_REQUIRE(!xs || (Bp >= B && ldx >= 3 * N_FEATS + 8 && (ldx & 7) == 0), AVHOT_ERR_INVALID,
                "frontend: split operand needs Bp >= B and a pitch >= 320 that is a multiple of 8");
  if (B == 0) return AVHOT_OK;
  const FrontendTables* tab = nullptr;
  int rc = get_tables(&tab);
  if (rc != AVHOT_OK) return rc;
  const int n_frames = 1 + n_samples / HOP;
  cudaStream_t s = as_stream(stream);
  if (xs && Bp > B) {   // rows of the padding clips are zero: one strided memset over the (T, Bp - B) tail of every step
    AVHOT_CUDA(cudaMemset2DAsync(reinterpret_cast<__half*>(xs) + (size_t)B * ldx, (size_t)Bp * ldx * sizeof(__half), 0,
                                 (size_t)(Bp - B) * ldx * sizeof(__half), (size_t)n_frames, s));
  }
  {
    // tensor map of the operand (3-D, so a 128-frame box that runs past the clip's last frame is clipped):
    // (n_frames, Bp, ldx) fp16 as (col, clip, t)
    CUtensorMap tmX;
    memset(&tmX, 0, sizeof(tmX));
    AVHOT_REQUIRE(!feats || (reinterpret_cast<uintptr_t>(feats) & 15) == 0, AVHOT_ERR_INVALID,
                  "frontend: feats must be 16-byte aligned");
    if (xs) {
      const uint64_t dims[3] = {(uint64_t)ldx, (uint64_t)Bp, (uint64_t)n_frames};
      const uint64_t strides[2] = {(uint64_t)ldx * 2, (uint64_t)Bp * ldx * 2};
      const uint32_t box[3] = {64, 1, FT_TILE};
      rc = tc::make_tensor_map_3d(&tmX, xs, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dims, strides, box);
      if (rc != AVHOT_OK) return rc;
    }
    const int smem = (int)sizeof(FtSmem) + 1024;
    AVHOT_CUDA(cudaFuncSetAttribute(frontend_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ProfScope prof("frontend", s);
    frontend_tc_kernel<<<min(B, num_sms()), FT_THREADS, smem, s>>>(wave, feats, tmX, xs ? 1 : 0, B, n_samples, n_frames, tab);
    AVHOT_LAUNCH_CHECK();
    return AVHOT_OK;
  }
}

extern "C" int avhot_frontend(const float* wave, float* feats, int B, int n_samples, avhot_stream_t stream) {
  AVHOT_REQUIRE(feats, AVHOT_ERR_INVALID, "frontend: feats is NULL");
  return frontend_launch(wave, feats, nullptr, 0, 0, B, n_samples, stream);
}

extern "C" int avhot_frontend_ex(const float* wave, float* feats, void* seq_split_f16, int B, int Bp, int pitch,
                                 int n_samples, avhot_stream_t stream) {
  return frontend_launch(wave, feats, seq_split_f16, Bp, pitch, B, n_samples, stream);
}

extern "C" int avhot_frontend_tables_host(float* mel_host, float* dct_host, float* window_host) {
  std::call_once(g_tables_once, build_tables);
  if (mel_host) memcpy(mel_host, g_host_mel_dense.data(), sizeof(float) * N_MELS * N_BINS);
  if (dct_host) memcpy(dct_host, g_host_dct.data(), sizeof(float) * N_MFCC * N_MELS);
  if (window_host) memcpy(window_host, g_host_tables.window, sizeof(float) * N_FFT);
  return AVHOT_OK;
}
