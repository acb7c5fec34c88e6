// fp32 FMA issue-rate microbenchmark (sm_100a): independent FFMA / FFMA2 chains in registers, no memory traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu && ./fma_peak
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int NACC>
__global__ void k(float* out, int iters, float a, float b) {
  float2 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x + i, i);
  const float2 a2 = make_float2(a, a + 1e-7f), b2 = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a2.x, b2.x); acc[i].y = fmaf(acc[i].y, a2.y, b2.y); }
        else acc[i] = __ffma2_rn(acc[i], a2, b2);
      }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int NACC> void run(const char* name, int threads) {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 4096, blocks = 148 * (2048 / threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, NACC><<<blocks, threads>>>(out, 16, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE, NACC><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * 2 * NACC * 8.0 * iters * (double)blocks * threads;
  printf("%-28s threads/CTA %4d acc pairs %2d: %.1f TFLOP/s\n", name, threads, NACC, flops / ms / 1e9);
  cudaFree(out);
}
int main() {
  run<0, 8>("FFMA", 1024); run<0, 16>("FFMA", 1024); run<0, 16>("FFMA", 256); run<0, 16>("FFMA", 128);
  run<1, 8>("FFMA2", 1024); run<1, 16>("FFMA2", 1024); run<1, 16>("FFMA2", 256); run<1, 16>("FFMA2", 128);
  return 0;
}
