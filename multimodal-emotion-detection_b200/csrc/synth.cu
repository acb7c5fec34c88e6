// Synthetic RAVDESS-shaped inputs, audio conditioning, and the Philox mask diagnostics.
// Written spec: oracle/synth.py, oracle/philox.py.  Everything here is bit-exact with the
// numpy oracle: only exact int->float conversions, power-of-two scalings and single
// correctly-rounded fp32 operations (explicit *_rn intrinsics so nothing is contracted
// into an FMA).
#include "common.cuh"

namespace avhot {

constexpr uint32_t STREAM_AUDIO_NOISE = 0x100, STREAM_AUDIO_ENV = 0x101, STREAM_AUDIO_TAIL = 0x102,
                   STREAM_VIDEO = 0x110;
constexpr int ENV_BLOCK = 800;

__device__ __forceinline__ float block_max_1024(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.f;
  r = warp_max(r);
  __syncthreads();
  return r;  // every thread of every warp holds the block max
}

// One CTA per clip.  Pass 1 writes the un-normalised samples and reduces max|y|;
// pass 2 divides in place (the clip is 192 KB: it is still in L2).
__global__ void __launch_bounds__(1024) synth_audio_kernel(float* __restrict__ wave, int n_samples,
                                                           uint64_t clip0, uint32_t k0, uint32_t k1) {
  __shared__ float red[32];
  const uint32_t g = (uint32_t)(clip0 + blockIdx.x);
  float* y = wave + (size_t)blockIdx.x * n_samples;
  const uint32_t tail = philox4x32_10(0, g, 0, STREAM_AUDIO_TAIL, k0, k1).x % 16001u;
  const int n_live = n_samples - (int)min(tail, (uint32_t)n_samples);
  float peak = 0.f;
  for (int i = threadIdx.x; i < n_samples; i += blockDim.x) {
    float v = 0.f;
    if (i < n_live) {
      const Philox4 w = philox4x32_10((uint32_t)i, g, 0, STREAM_AUDIO_NOISE, k0, k1);
      const uint32_t s = (w.x >> 10) + (w.y >> 10) + (w.z >> 10) + (w.w >> 10);
      const float noise = __fmul_rn(__fsub_rn((float)s, 8388606.0f), 2.384185791015625e-07f);
      const uint32_t b = (uint32_t)i / ENV_BLOCK;
      const uint32_t ew = philox_word(philox4x32_10(b >> 2, g, 0, STREAM_AUDIO_ENV, k0, k1), b & 3);
      const float u = __fmul_rn((float)(ew >> 8), 5.9604644775390625e-08f);
      const float gain = __fadd_rn(0.05f, __fmul_rn(0.95f, u));
      v = __fmul_rn(noise, gain);
    }
    y[i] = v;
    peak = fmaxf(peak, fabsf(v));
  }
  peak = block_max_1024(peak, red);
  if (peak > 0.f) {
    for (int i = threadIdx.x; i < n_samples; i += blockDim.x) y[i] = __fdiv_rn(y[i], peak);
  }
}

__global__ void synth_video_kernel(float* __restrict__ f32, uint8_t* __restrict__ u8,
                                   int n_per_clip, uint64_t clip0, uint32_t k0, uint32_t k1) {
  const uint32_t g = (uint32_t)(clip0 + blockIdx.y);
  const int nblk = (n_per_clip + 3) >> 2;
  for (int blk = blockIdx.x * blockDim.x + threadIdx.x; blk < nblk; blk += gridDim.x * blockDim.x) {
    const Philox4 w = philox4x32_10((uint32_t)blk, g, 0, STREAM_VIDEO, k0, k1);
    const uint32_t px[4] = {w.x >> 24, w.y >> 24, w.z >> 24, w.w >> 24};
    const size_t base = (size_t)blockIdx.y * n_per_clip + (size_t)blk * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (blk * 4 + j < n_per_clip) {
        if (f32) f32[base + j] = __fdiv_rn((float)px[j], 255.0f);
        if (u8) u8[base + j] = (uint8_t)px[j];
      }
    }
  }
}

// a1: src/dataprocessing.py:198-211
__global__ void __launch_bounds__(1024) condition_audio_kernel(const float* __restrict__ in,
                                                               const int32_t* __restrict__ lengths,
                                                               float* __restrict__ out, int in_stride,
                                                               int n_out) {
  __shared__ float red[32];
  const float* x = in + (size_t)blockIdx.x * in_stride;
  float* y = out + (size_t)blockIdx.x * n_out;
  int len = lengths ? lengths[blockIdx.x] : in_stride;
  len = max(0, min(len, min(in_stride, n_out)));        // truncate (:199-200)
  float peak = 0.f;
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
    const float v = i < len ? x[i] : 0.f;                // right zero pad (:201-204)
    y[i] = v;
    peak = fmaxf(peak, fabsf(v));
  }
  peak = block_max_1024(peak, red);
  if (peak > 0.f) {                                      // :207-208
    for (int i = threadIdx.x; i < n_out; i += blockDim.x) y[i] = __fdiv_rn(y[i], peak);
  }
}

// the same for 16-bit PCM (what a RAVDESS .wav holds): sample k means k / 32768 (soundfile / librosa.load's conversion,
// src/dataprocessing.py:196); (k / 32768) / (peak / 32768) == fp32(k) / fp32(peak) exactly, both being the correctly
// rounded quotient of the same real number, so the result is bit-identical to the fp32 entry point's
__global__ void __launch_bounds__(1024) condition_audio_i16_kernel(const int16_t* __restrict__ in,
                                                                   const int32_t* __restrict__ lengths,
                                                                   float* __restrict__ out, int in_stride, int n_out) {
  __shared__ float red[32];
  const int16_t* x = in + (size_t)blockIdx.x * in_stride;
  float* y = out + (size_t)blockIdx.x * n_out;
  int len = lengths ? lengths[blockIdx.x] : in_stride;
  len = max(0, min(len, min(in_stride, n_out)));
  float peak = 0.f;
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) peak = fmaxf(peak, i < len ? fabsf((float)x[i]) : 0.f);
  peak = block_max_1024(peak, red);
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
    const float v = i < len ? (float)x[i] : 0.f;
    y[i] = peak > 0.f ? __fdiv_rn(v, peak) : 0.f;
  }
}

__global__ void dropout_keep_kernel(uint8_t* __restrict__ keep, int n_elems, uint64_t clip0, uint32_t t,
                                    uint32_t site, uint32_t thr, uint32_t k0, uint32_t k1) {
  const uint32_t g = (uint32_t)(clip0 + blockIdx.y);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += gridDim.x * blockDim.x) {
    bool k;
    if (n_elems >= AVHOT_DENSE_SITE_ELEMS)
      k = philox_half(philox4x32_10((uint32_t)e >> 3, g, t, site, k0, k1), e & 7) >= (thr >> 16);
    else
      k = philox_word(philox4x32_10((uint32_t)e >> 2, g, t, site, k0, k1), e & 3) >= thr;
    keep[(size_t)blockIdx.y * n_elems + e] = k ? 1 : 0;
  }
}

// src/data.py:108-118
__global__ void modality_mask_kernel(float* __restrict__ mask, int B, int M, uint64_t clip0, uint32_t thr,
                                     int active, uint32_t k0, uint32_t k1) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const Philox4 w = philox4x32_10(0, (uint32_t)(clip0 + b), 0, AVHOT_SITE_MODMASK, k0, k1);
  bool any = false;
  bool keep[3];
  for (int m = 0; m < M; ++m) {
    keep[m] = !active || philox_word(w, m) > thr;
    any |= keep[m];
  }
  if (!any) keep[w.w % (uint32_t)M] = true;
  for (int m = 0; m < M; ++m) mask[(size_t)b * M + m] = keep[m] ? 1.f : 0.f;
}

}  // namespace avhot

using namespace avhot;

extern "C" int avhot_synth_audio(float* wave, int B, int n_samples, uint64_t clip0, uint64_t seed,
                                 avhot_stream_t stream) {
  AVHOT_REQUIRE(wave && B >= 0 && n_samples > 0, AVHOT_ERR_INVALID, "synth_audio: bad arguments");
  if (B == 0) return AVHOT_OK;
  synth_audio_kernel<<<B, 1024, 0, as_stream(stream)>>>(wave, n_samples, clip0, (uint32_t)seed,
                                                        (uint32_t)(seed >> 32));
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

extern "C" int avhot_synth_video(float* frames_f32, uint8_t* frames_u8, int B, int n_elems_per_clip,
                                 uint64_t clip0, uint64_t seed, avhot_stream_t stream) {
  AVHOT_REQUIRE((frames_f32 || frames_u8) && B >= 0 && n_elems_per_clip > 0, AVHOT_ERR_INVALID,
                "synth_video: bad arguments");
  AVHOT_REQUIRE(B <= 65535, AVHOT_ERR_INVALID, "synth_video: B > 65535 per call");
  if (B == 0) return AVHOT_OK;
  dim3 grid(min(ceil_div(ceil_div(n_elems_per_clip, 4), 256), 64), B);
  synth_video_kernel<<<grid, 256, 0, as_stream(stream)>>>(frames_f32, frames_u8, n_elems_per_clip, clip0,
                                                          (uint32_t)seed, (uint32_t)(seed >> 32));
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

extern "C" int avhot_condition_audio(const float* in, const int32_t* lengths, float* out, int B,
                                     int in_stride, int n_out, avhot_stream_t stream) {
  AVHOT_REQUIRE(in && out && B >= 0 && in_stride > 0 && n_out > 0, AVHOT_ERR_INVALID,
                "condition_audio: bad arguments");
  if (B == 0) return AVHOT_OK;
  condition_audio_kernel<<<B, 1024, 0, as_stream(stream)>>>(in, lengths, out, in_stride, n_out);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

extern "C" int avhot_condition_audio_i16(const int16_t* in, const int32_t* lengths, float* out, int B,
                                         int in_stride, int n_out, avhot_stream_t stream) {
  AVHOT_REQUIRE(in && out && B >= 0 && in_stride > 0 && n_out > 0, AVHOT_ERR_INVALID,
                "condition_audio_i16: bad arguments");
  if (B == 0) return AVHOT_OK;
  ProfScope prof("condition_audio_i16", as_stream(stream));
  condition_audio_i16_kernel<<<B, 1024, 0, as_stream(stream)>>>(in, lengths, out, in_stride, n_out);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

extern "C" int avhot_dropout_keep(uint8_t* keep, int B, int n_elems, uint64_t clip0, uint32_t t,
                                  uint32_t site, float p, uint64_t seed, avhot_stream_t stream) {
  AVHOT_REQUIRE(keep && B >= 0 && n_elems > 0 && p >= 0.f && p < 1.f, AVHOT_ERR_INVALID,
                "dropout_keep: bad arguments");
  AVHOT_REQUIRE(B <= 65535, AVHOT_ERR_INVALID, "dropout_keep: B > 65535 per call");
  if (B == 0) return AVHOT_OK;
  dim3 grid(min(ceil_div(n_elems, 256), 256), B);
  dropout_keep_kernel<<<grid, 256, 0, as_stream(stream)>>>(keep, n_elems, clip0, t, site,
                                                           drop_threshold(p), (uint32_t)seed,
                                                           (uint32_t)(seed >> 32));
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

extern "C" int avhot_modality_mask(float* mask, int B, int M, uint64_t clip0, float p, uint64_t seed,
                                   avhot_stream_t stream) {
  AVHOT_REQUIRE(mask && B >= 0 && M >= 1 && M <= 3 && p >= 0.f && p < 1.f, AVHOT_ERR_INVALID,
                "modality_mask: bad arguments (1 <= M <= 3)");
  if (B == 0) return AVHOT_OK;
  modality_mask_kernel<<<ceil_div(B, 256), 256, 0, as_stream(stream)>>>(
      mask, B, M, clip0, drop_threshold(p), p > 0.f ? 1 : 0, (uint32_t)seed, (uint32_t)(seed >> 32));
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}
