// Label exactness of the tensor-core path (BASELINE north star: "predicted labels identical to the reference").
//
// argmax(mean_logits) (src/uncertainty.py:77 feeds the prediction) of the tensor-core path can differ from the fp32
// path's only for clips whose top-2 margin is below the error of that difference (measured envelope: exact.py).  Those
// clips - under one percent - are appended to a pending list with a copy of their inputs, re-evaluated through the fp32
// kernels (each with its own global clip index, so the Philox masks are the ones it drew the first time) and their
// records overwritten:
//   avhot_flag_near_ties    top-2 margins on the device, list append by atomics (no host round trip);
//   avhot_gather_flagged    rows of an input tensor -> the pending pool, for the entries the last flag call appended;
//   avhot_scatter_resolved  fp32 records of the pending entries of one call -> that call's result tensors.
#include "common.cuh"

namespace avhot {

// state words: [0] entries appended (may exceed capacity), [1] overflow flag, [2] entries before the latest flag call
__global__ void near_tie_flag_kernel(const float* __restrict__ mean_logits, int B, int C, float margin, uint32_t tag,
                                     uint32_t clip0_lo, uint32_t* __restrict__ state, uint32_t* __restrict__ entry_tag,
                                     uint32_t* __restrict__ entry_row, uint32_t* __restrict__ entry_clip, int capacity) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* x = mean_logits + (size_t)b * C;
  float top1 = -INFINITY, top2 = -INFINITY;
  bool nan = false;
  for (int c = 0; c < C; ++c) {
    const float v = x[c];
    nan |= v != v;
    if (v > top1) { top2 = top1; top1 = v; }
    else if (v > top2) top2 = v;
  }
  // a row with a NaN logit is flagged too, so the fp32 path has the last word on it
  if (C >= 2 && (nan || !(top1 - top2 >= margin))) {
    const uint32_t e = atomicAdd(&state[0], 1u);
    if (e < (uint32_t)capacity) {
      entry_tag[e] = tag;
      entry_row[e] = (uint32_t)b;
      entry_clip[e] = clip0_lo + (uint32_t)b;
    } else {
      state[1] = 1u;
    }
  }
}

__global__ void near_tie_begin_kernel(uint32_t* __restrict__ state) { state[2] = state[0]; }

// CTA i copies the row of the i-th entry appended by the latest flag call, in words of W bytes (uint4 when rows and
// bases are 16-byte multiples, else 32-bit words: the (B, 2) modality mask has 8-byte rows)
template <typename W>
__global__ void gather_flagged_kernel(const W* __restrict__ src, size_t row_words, const uint32_t* __restrict__ state,
                                      const uint32_t* __restrict__ entry_row, W* __restrict__ pool, int capacity) {
  const uint32_t first = state[2], last = min(state[0], (uint32_t)capacity);
  for (uint32_t e = first + blockIdx.x; e < last; e += gridDim.x) {
    const W* s = src + (size_t)entry_row[e] * row_words;
    W* d = pool + (size_t)e * row_words;
    for (size_t i = threadIdx.x; i < row_words; i += blockDim.x) d[i] = __ldg(s + i);
  }
}

__global__ void scatter_resolved_kernel(const uint32_t* __restrict__ entry_tag, const uint32_t* __restrict__ entry_row,
                                        int n, uint32_t tag, const float* __restrict__ src_packed,
                                        const int32_t* __restrict__ src_labels, float* __restrict__ mean_logits,
                                        float* __restrict__ uncertainty, int32_t* __restrict__ labels,
                                        float* __restrict__ packed, int C) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || entry_tag[e] != tag) return;
  const size_t row = entry_row[e];
  const float* s = src_packed + (size_t)e * (C + 1);
  for (int c = 0; c < C; ++c) {
    if (mean_logits) mean_logits[row * C + c] = s[c];
    if (packed) packed[row * (C + 1) + c] = s[c];
  }
  if (uncertainty) uncertainty[row] = s[C];
  if (packed) packed[row * (C + 1) + C] = s[C];
  if (labels) labels[row] = src_labels[e];
}

}  // namespace avhot

using namespace avhot;

extern "C" int avhot_flag_near_ties(const float* mean_logits, int B, int C, float margin, uint32_t tag,
                                    uint64_t clip0, uint32_t* state, uint32_t* entry_tag, uint32_t* entry_row,
                                    uint32_t* entry_clip, int capacity, avhot_stream_t stream) {
  AVHOT_REQUIRE(mean_logits && state && entry_tag && entry_row && entry_clip && B >= 0 && C >= 1 && capacity >= 1,
                AVHOT_ERR_INVALID, "flag_near_ties: bad arguments");
  AVHOT_REQUIRE(margin >= 0.f, AVHOT_ERR_INVALID, "flag_near_ties: margin=%g < 0", (double)margin);
  if (B == 0) return AVHOT_OK;
  cudaStream_t s = as_stream(stream);
  ProfScope prof("near_tie_flag", s);
  near_tie_begin_kernel<<<1, 1, 0, s>>>(state);
  AVHOT_LAUNCH_CHECK();
  near_tie_flag_kernel<<<ceil_div(B, 256), 256, 0, s>>>(mean_logits, B, C, margin, tag, (uint32_t)clip0, state,
                                                       entry_tag, entry_row, entry_clip, capacity);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

extern "C" int avhot_gather_flagged(const void* src, size_t row_bytes, int B, const uint32_t* state,
                                    const uint32_t* entry_row, void* pool, int capacity, avhot_stream_t stream) {
  AVHOT_REQUIRE(src && state && entry_row && pool && B >= 0 && capacity >= 1, AVHOT_ERR_INVALID,
                "gather_flagged: bad arguments");
  AVHOT_REQUIRE(row_bytes > 0 && (row_bytes & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(pool) & 3) == 0,
                AVHOT_ERR_INVALID, "gather_flagged: rows must be 4-byte multiples, 4-byte aligned");
  if (B == 0) return AVHOT_OK;
  cudaStream_t s = as_stream(stream);
  ProfScope prof("near_tie_gather", s);
  // a flag call appends at most B entries; a few hundred CTAs walk them (most calls flag a fraction of a percent)
  const int grid = min(min(B, capacity), 4 * num_sms());
  if ((row_bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(pool) & 15) == 0)
    gather_flagged_kernel<uint4><<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(src), row_bytes / 16, state, entry_row,
                                                     reinterpret_cast<uint4*>(pool), capacity);
  else
    gather_flagged_kernel<uint32_t><<<grid, 32, 0, s>>>(reinterpret_cast<const uint32_t*>(src), row_bytes / 4, state,
                                                       entry_row, reinterpret_cast<uint32_t*>(pool), capacity);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

extern "C" int avhot_scatter_resolved(const uint32_t* entry_tag, const uint32_t* entry_row, int n, uint32_t tag,
                                      const float* src_packed, const int32_t* src_labels, float* mean_logits,
                                      float* uncertainty, int32_t* labels, float* packed, int C,
                                      avhot_stream_t stream) {
  AVHOT_REQUIRE(entry_tag && entry_row && src_packed && src_labels && n >= 0 && C >= 1, AVHOT_ERR_INVALID,
                "scatter_resolved: bad arguments");
  if (n == 0) return AVHOT_OK;
  cudaStream_t s = as_stream(stream);
  ProfScope prof("near_tie_scatter", s);
  scatter_resolved_kernel<<<ceil_div(n, 128), 128, 0, s>>>(entry_tag, entry_row, n, tag, src_packed, src_labels,
                                                          mean_logits, uncertainty, labels, packed, C);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}
