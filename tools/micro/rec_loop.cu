// The fp32 LSTM recurrence's inner loop in isolation (sm_100a): 512 threads = (unit j, clip half), per k-row one 128-bit
// weight load (per-lane), 6 broadcast h values, 12 FFMA2 - with the pieces switched on one at a time, to see what holds
// lstm_recurrence_kernel at half the FMA rate.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rec_loop rec_loop.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int LH = 256, ROWS = 256;
template <bool DO_W, bool DO_H, bool DO_FMA, int SYNC_EVERY>
__global__ void __launch_bounds__(512, 1) k(float* out, int steps) {
  extern __shared__ __align__(16) unsigned char raw[];
  float4* w = reinterpret_cast<float4*>(raw);                       // [32 rows][256 units] ring-like window (reused)
  float (*hs)[16] = reinterpret_cast<float (*)[16]>(raw + 32 * LH * 16);   // [256 k][16]
  const int j = threadIdx.x & 255, half = threadIdx.x >> 8;
  for (int i = threadIdx.x; i < 32 * LH; i += 512) w[i] = make_float4(1e-3f * i, 1.f, 2.f, 3.f);
  for (int i = threadIdx.x; i < 256 * 16; i += 512) (&hs[0][0])[i] = 1e-4f * i;
  __syncthreads();
  float2 acc[4][3];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int b = 0; b < 3; ++b) acc[g][b] = make_float2(0.f, 0.f);
  float4 w4 = make_float4(1.f, 2.f, 3.f, 4.f);
  float2 hv[3] = {make_float2(1.f, 2.f), make_float2(3.f, 4.f), make_float2(5.f, 6.f)};
  for (int s = 0; s < steps; ++s) {
#pragma unroll 8
    for (int kk = 0; kk < ROWS; ++kk) {
      if (SYNC_EVERY && (kk % SYNC_EVERY) == 0) __syncthreads();
      if (DO_W) w4 = w[(kk & 31) * LH + j];
      if (DO_H) {
        const float4 v = *reinterpret_cast<const float4*>(&hs[kk][half * 8]);
        const float2 v2 = *reinterpret_cast<const float2*>(&hs[kk][half * 8 + 4]);
        hv[0] = make_float2(v.x, v.y); hv[1] = make_float2(v.z, v.w); hv[2] = v2;
      }
      if (DO_FMA) {
        const float2 wg[4] = {make_float2(w4.x, w4.x), make_float2(w4.y, w4.y), make_float2(w4.z, w4.z), make_float2(w4.w, w4.w)};
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int b = 0; b < 3; ++b) acc[g][b] = __ffma2_rn(wg[g], hv[b], acc[g][b]);
      } else {
        acc[0][0].x += w4.x + w4.y + w4.z + w4.w + hv[0].x + hv[1].x + hv[2].y;
      }
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int b = 0; b < 3; ++b) sum += acc[g][b].x + acc[g][b].y;
  out[blockIdx.x * 512 + threadIdx.x] = sum;
}
// 256 threads = one per unit, all 12 clips of the tile per thread (24 FFMA2 per row from 4 + 12 loaded floats: 56 shared-
// memory wavefronts per row instead of 96), operands fetched PF rows ahead into rotating registers
template <int PF, int SYNC_EVERY>
__global__ void __launch_bounds__(256, 1) k256(float* out, int steps) {
  extern __shared__ __align__(16) unsigned char raw[];
  float4* w = reinterpret_cast<float4*>(raw);
  float (*hs)[16] = reinterpret_cast<float (*)[16]>(raw + 32 * LH * 16);
  const int j = threadIdx.x;
  for (int i = threadIdx.x; i < 32 * LH; i += 256) w[i] = make_float4(1e-3f * i, 1.f, 2.f, 3.f);
  for (int i = threadIdx.x; i < 256 * 16; i += 256) (&hs[0][0])[i] = 1e-4f * i;
  __syncthreads();
  float2 acc[4][6];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int b = 0; b < 6; ++b) acc[g][b] = make_float2(0.f, 0.f);
  for (int s = 0; s < steps; ++s) {
    for (int k0 = 0; k0 < ROWS; k0 += 16) {
      if (SYNC_EVERY) __syncthreads();
      float4 wq[PF + 1];
      float4 hq[PF + 1][3];
#pragma unroll
      for (int p = 0; p < PF; ++p) {
        wq[p] = w[((k0 + p) & 31) * LH + j];
#pragma unroll
        for (int q = 0; q < 3; ++q) hq[p][q] = *reinterpret_cast<const float4*>(&hs[k0 + p][4 * q]);
      }
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        if (kk + PF < 16) {
          wq[(kk + PF) % (PF + 1)] = w[((k0 + kk + PF) & 31) * LH + j];
#pragma unroll
          for (int q = 0; q < 3; ++q) hq[(kk + PF) % (PF + 1)][q] = *reinterpret_cast<const float4*>(&hs[k0 + kk + PF][4 * q]);
        }
        const float4 w4 = wq[kk % (PF + 1)];
        const float2 wg[4] = {make_float2(w4.x, w4.x), make_float2(w4.y, w4.y), make_float2(w4.z, w4.z), make_float2(w4.w, w4.w)};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const float4 h4 = hq[kk % (PF + 1)][q];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            acc[g][2 * q] = __ffma2_rn(wg[g], make_float2(h4.x, h4.y), acc[g][2 * q]);
            acc[g][2 * q + 1] = __ffma2_rn(wg[g], make_float2(h4.z, h4.w), acc[g][2 * q + 1]);
          }
        }
      }
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int b = 0; b < 6; ++b) sum += acc[g][b].x + acc[g][b].y;
  out[blockIdx.x * 256 + threadIdx.x] = sum;
}
template <int PF, int S> void run256(const char* name) {
  float* out; cudaMalloc(&out, 148 * 256 * 4);
  auto kern = k256<PF, S>;
  const int smem = 32 * LH * 16 + 256 * 16 * 4;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int steps = 600;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148, 256, smem>>>(out, 4);
  cudaEventRecord(e0);
  kern<<<148, 256, smem>>>(out, steps);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-44s %.2f us per step of 256 rows (%.0f cycles per row at 1.9 GHz)\n", name, ms * 1e3 / steps, ms * 1e-3 / steps / 256 * 1.9e9);
}
template <bool W, bool H, bool F, int S> void run(const char* name) {
  float* out; cudaMalloc(&out, 148 * 512 * 4);
  auto kern = k<W, H, F, S>;
  const int smem = 32 * LH * 16 + 256 * 16 * 4;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int steps = 600;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148, 512, smem>>>(out, 4);
  cudaEventRecord(e0);
  kern<<<148, 512, smem>>>(out, steps);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-44s %.2f us per step of 256 rows (%.0f cycles per row at 1.9 GHz)\n", name, ms * 1e3 / steps, ms * 1e-3 / steps / 256 * 1.9e9);
}
int main() {
  run<true, true, true, 0>("weights + h + FFMA2");
  run<true, true, true, 8>("weights + h + FFMA2, barrier every 8 rows");
  run<false, false, true, 0>("FFMA2 only");
  run<true, false, true, 0>("weights + FFMA2");
  run<false, true, true, 0>("h + FFMA2");
  run<true, true, false, 0>("loads only");
  run256<1, 0>("256 threads x 12 clips, 1 row ahead");
  run256<2, 0>("256 threads x 12 clips, 2 rows ahead");
  run256<3, 0>("256 threads x 12 clips, 3 rows ahead");
  run256<2, 1>("256 threads, 2 ahead, barrier every 16 rows");
  return 0;
}
