// Register-only outer-product FMA stream shaped like the fp32 LSTM recurrence's inner loop (4 gate weights x 12 clips per
// k-row), operands rotated through registers so nothing is loop-invariant: what rate does the FFMA / FFMA2 stream reach
// without any shared-memory traffic?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_outer fma_outer.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, const float* in, int iters) {
  float2 acc[4][6];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int b = 0; b < 6; ++b) acc[g][b] = make_float2(0.f, 0.f);
  float w[4], h[12];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = in[threadIdx.x + i];
#pragma unroll
  for (int i = 0; i < 12; ++i) h[i] = in[threadIdx.x + 4 + i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int b = 0; b < 6; ++b) {
          if (MODE == 0) { acc[g][b].x = fmaf(w[g], h[2 * b], acc[g][b].x); acc[g][b].y = fmaf(w[g], h[2 * b + 1], acc[g][b].y); }
          else acc[g][b] = __ffma2_rn(make_float2(w[g], w[g]), make_float2(h[2 * b], h[2 * b + 1]), acc[g][b]);
        }
      // new operands every row (cheap integer-free rotation so the compiler cannot hoist them)
      const float t = w[0]; w[0] = w[1]; w[1] = w[2]; w[2] = w[3]; w[3] = h[0];
#pragma unroll
      for (int i = 0; i < 11; ++i) h[i] = h[i + 1];
      h[11] = t;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int b = 0; b < 6; ++b) s += acc[g][b].x + acc[g][b].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name) {
  float *out, *in; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
  const int iters = 8192;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, 512>>>(out, in, 16);
  cudaEventRecord(e0);
  k<MODE><<<148, 512>>>(out, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-8s 4 x 12 outer product per row, 512 threads/SM: %.1f TFLOP/s\n", name, 2.0 * 48 * 8.0 * iters * 148 * 512 / ms / 1e9);
}
int main() { run<0>("FFMA"); run<1>("FFMA2"); return 0; }
