// fp32 CUDA-core GEMM:  C[M,N] = act(A[M,K] * W[N,K]^T + bias[N])   (AVHOT_PREC_FP32 path)
//
// Both operands are K-contiguous (activations row-major, nn.Linear weights (out, in)), so a
// 128 x 128 x 16 CTA tile is loaded with 128-bit loads, transposed into shared memory as
// [k][m] / [k][n], and each of the 256 threads accumulates an 8 x 8 register tile from
// conflict-free 128-bit shared loads.  The next k-slab is prefetched into registers while the
// current one is multiplied.  A may also be u8 pixels (k -> fp32(k)/255, the reference's
// frame normalisation, src/dataprocessing.py:261-265), converted on load.
#include "common.cuh"

namespace avhot {

constexpr int GBM = 128, GBN = 128, GBK = 16, GTHREADS = 256;

template <bool A_U8>
__global__ void __launch_bounds__(GTHREADS)
gemm_f32_kernel(const float* __restrict__ A, const uint8_t* __restrict__ A8, int lda,
                const float* __restrict__ W, int ldw, const float* __restrict__ bias,
                float* __restrict__ C, int ldc, int M, int N, int K, int act) {
  __shared__ __align__(16) float As[2][GBK][GBM + 4];
  __shared__ __align__(16) float Ws[2][GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  // loader mapping: 128 rows x 16 k = 512 float4; thread loads (row = tid >> 2 [+64], k4 = tid & 3)
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float4 ra[2], rw[2];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + 64 * h;
      const int k = k0 + lk;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vw = va;
      if (m0 + r < M && k < K) {
        if (A_U8) {
          const uchar4 p = *reinterpret_cast<const uchar4*>(A8 + (size_t)(m0 + r) * lda + k);
          va = make_float4(__fdiv_rn((float)p.x, 255.f), __fdiv_rn((float)p.y, 255.f),
                           __fdiv_rn((float)p.z, 255.f), __fdiv_rn((float)p.w, 255.f));
        } else {
          va = __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * lda + k));
        }
      }
      if (n0 + r < N && k < K) vw = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + r) * ldw + k));
      ra[h] = va;
      rw[h] = vw;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + 64 * h;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y;
      As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Ws[buf][lk + 0][r] = rw[h].x; Ws[buf][lk + 1][r] = rw[h].y;
      Ws[buf][lk + 2][r] = rw[h].z; Ws[buf][lk + 3][r] = rw[h].w;
    }
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (K + GBK - 1) / GBK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * GBK);
    // fragments of row k + 1 are fetched into a second register set before the FMAs of row k (ptxas otherwise leaves the
    // shared loads a few instructions in front of their first use: tools/micro/rec_loop.cu)
    float4 fa[2][2], fb[2][2];
    fa[0][0] = *reinterpret_cast<const float4*>(&As[buf][0][ty * 4]);
    fa[0][1] = *reinterpret_cast<const float4*>(&As[buf][0][64 + ty * 4]);
    fb[0][0] = *reinterpret_cast<const float4*>(&Ws[buf][0][tx * 4]);
    fb[0][1] = *reinterpret_cast<const float4*>(&Ws[buf][0][64 + tx * 4]);
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      if (k + 1 < GBK) {
        fa[(k + 1) & 1][0] = *reinterpret_cast<const float4*>(&As[buf][k + 1][ty * 4]);
        fa[(k + 1) & 1][1] = *reinterpret_cast<const float4*>(&As[buf][k + 1][64 + ty * 4]);
        fb[(k + 1) & 1][0] = *reinterpret_cast<const float4*>(&Ws[buf][k + 1][tx * 4]);
        fb[(k + 1) & 1][1] = *reinterpret_cast<const float4*>(&Ws[buf][k + 1][64 + tx * 4]);
      }
      const float4 a0 = fa[k & 1][0], a1 = fa[k & 1][1], b0 = fb[k & 1][0], b1 = fb[k & 1][1];
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: bias + activation
  const bool vec_ok = (ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][jh * 4 + j];
        if (bias && n + j < N) x += __ldg(bias + n + j);
        if (act == ACT_RELU) x = fmaxf(x, 0.f);
        v[j] = x;
      }
      float* dst = C + (size_t)m * ldc + n;
      if (vec_ok && n + 3 < N) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) dst[j] = v[j];
      }
    }
  }
}

int gemm_f32(const float* a_f32, const uint8_t* a_u8, int lda, const float* w, int ldw,
             const float* bias, float* c, int ldc, int M, int N, int K, int act, cudaStream_t s,
             const char* prof_name) {
  AVHOT_REQUIRE((a_f32 != nullptr) != (a_u8 != nullptr), AVHOT_ERR_INVALID, "gemm: exactly one A operand");
  AVHOT_REQUIRE(w && c && M >= 0 && N > 0 && K > 0, AVHOT_ERR_INVALID, "gemm: bad arguments");
  AVHOT_REQUIRE((K & 3) == 0 && (lda & 3) == 0 && (ldw & 3) == 0, AVHOT_ERR_INVALID,
                "gemm: K, lda, ldw must be multiples of 4 (K=%d lda=%d ldw=%d)", K, lda, ldw);
  AVHOT_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                    (a_f32 ? (reinterpret_cast<uintptr_t>(a_f32) & 15) == 0
                           : (reinterpret_cast<uintptr_t>(a_u8) & 3) == 0),
                AVHOT_ERR_INVALID, "gemm: operands must be 16-byte aligned");
  if (M == 0) return AVHOT_OK;
  dim3 grid(ceil_div(N, GBN), ceil_div(M, GBM));
  AVHOT_REQUIRE(grid.y <= 65535, AVHOT_ERR_INVALID, "gemm: M too large for one launch");
  ProfScope prof(prof_name, s);
  if (a_u8)
    gemm_f32_kernel<true><<<grid, GTHREADS, 0, s>>>(nullptr, a_u8, lda, w, ldw, bias, c, ldc, M, N, K, act);
  else
    gemm_f32_kernel<false><<<grid, GTHREADS, 0, s>>>(a_f32, nullptr, lda, w, ldw, bias, c, ldc, M, N, K, act);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

}  // namespace avhot
