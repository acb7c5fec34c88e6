// Library plumbing: error strings, version, device check, small layout kernels.
#include <stdarg.h>

#include "common.cuh"

namespace avhot {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// dst[c][r] = src[r][c]  (32x32 tiles through shared memory, coalesced both ways)
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? src[(size_t)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

int transpose(const float* src, float* dst, int R, int C, cudaStream_t s) {
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, s>>>(src, dst, R, C);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b,
                               float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + (b ? b[i] : 0.f);
}

int add_vec(const float* a, const float* b, float* out, int n, cudaStream_t s) {
  add_vec_kernel<<<ceil_div(n, 256), 256, 0, s>>>(a, b, out, n);
  AVHOT_LAUNCH_CHECK();
  return AVHOT_OK;
}

}  // namespace avhot

extern "C" int avhot_version(void) { return AVHOT_VERSION; }

extern "C" const char* avhot_last_error(void) { return avhot::g_err; }

extern "C" int avhot_check_device(void) {
  int dev = 0;
  AVHOT_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  AVHOT_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  AVHOT_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  AVHOT_REQUIRE(major == 10 && minor == 0, AVHOT_ERR_UNSUPPORTED,
                "libavhot is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
  return AVHOT_OK;
}
