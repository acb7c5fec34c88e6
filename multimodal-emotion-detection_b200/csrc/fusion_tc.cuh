// Interface between fusion.cu (handle, dispatch) and fusion_tc.cu (tcgen05 per-pass kernel).
#pragma once
#include "common.cuh"

namespace avhot {

// small fp32 vectors the epilogue needs, packed into one array (copied to shared memory by the kernel)
enum {
  VEC_PROJ_B = 0,            // proj_layers.<video>.bias
  VEC_PRE_W = 256, VEC_PRE_B = 512, VEC_POST_W = 768, VEC_POST_B = 1024,
  VEC_BK0 = 1280, VEC_BV0 = 1536,                 // LN-folded biases, layer 0 (query = audio token)
  VEC_BQ1 = 1792, VEC_BK1 = 2048, VEC_BV1 = 2304, // layer 1 (query = video token)
  VEC_BO0 = 2560, VEC_BO1 = 2816,
  VEC_G0_B = 3072, VEC_G3_W = 3200, VEC_G3_B = 3328,
  VEC_CLS_W = 3332,          // (8, 256), rows beyond num_classes are zero
  VEC_CLS_B = 3332 + 2048,
  FUS_VEC_FLOATS = 3332 + 2048 + 8
};

struct FusTcParams {
  const float* z_v;       // (B * rpc_v, dim_v)
  int rpc_v, dim_v;
  const float* clipbuf;   // (B, 6, 256): Z_a, Q0, K0a, V0a, K1a, V1a  (fusion_clip_kernel)
  const float* mask;      // (B, 2) or NULL
  float* logits;          // (B * T, C)
  int B, T, C;
  const float* vec;       // FUS_VEC_FLOATS floats
  DropSpec ds;
  long long* trace;       // debug (AVHOT_FUS_TRACE=<file>): clock64() at the epilogue's phase boundaries, or NULL
};

size_t fusion_tc_packed_bytes();
// rows [row0, row0 + N) of the packed fp16 weight tensor from a k-major fp32 matrix src_t[k][n]
int fusion_tc_pack_rows(const float* src_t, void* packed, int row0, int N, int K, cudaStream_t s);
int fusion_tc_forward(const void* packed_w, const FusTcParams& P, cudaStream_t s);

}  // namespace avhot
