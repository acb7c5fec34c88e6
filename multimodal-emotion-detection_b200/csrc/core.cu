// Library plumbing: error strings, version, device check, small layout kernels.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace avhot {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfEntry {
  std::string name;
  std::vector<cudaEvent_t> ev;   // start, stop, start, stop, ...
};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mutex;
static std::vector<ProfEntry> g_prof;

ProfScope::ProfScope(const char* name, cudaStream_t stream) : s(stream), slot(-1) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  for (size_t i = 0; i < g_prof.size(); ++i)
    if (g_prof[i].name == name) slot = (int)i;
  if (slot < 0) {
    g_prof.push_back(ProfEntry{name, {}});
    slot = (int)g_prof.size() - 1;
  }
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) { slot = -1; return; }
  cudaEventRecord(e, s);
  g_prof[slot].ev.push_back(e);
}

ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) { g_prof[slot].ev.pop_back(); return; }
  cudaEventRecord(e, s);
  g_prof[slot].ev.push_back(e);
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

extern "C" long long avhot_launch_count(void) { return avhot::g_launches.load(); }

extern "C" int avhot_prof_enable(int on) {
  avhot::g_prof_on.store(on ? 1 : 0);
  return AVHOT_OK;
}

extern "C" int avhot_prof_reset(void) {
  std::lock_guard<std::mutex> lock(avhot::g_prof_mutex);
  for (auto& p : avhot::g_prof)
    for (cudaEvent_t e : p.ev) cudaEventDestroy(e);
  avhot::g_prof.clear();
  return AVHOT_OK;
}

extern "C" int avhot_prof_read(const char* name, double* total_ms, int* launches) {
  AVHOT_REQUIRE(name && total_ms && launches, AVHOT_ERR_INVALID, "prof_read: null argument");
  std::lock_guard<std::mutex> lock(avhot::g_prof_mutex);
  *total_ms = 0.0;
  *launches = 0;
  for (auto& p : avhot::g_prof) {
    if (p.name != name) continue;
    for (size_t i = 0; i + 1 < p.ev.size(); i += 2) {
      AVHOT_CUDA(cudaEventSynchronize(p.ev[i + 1]));
      float ms = 0.f;
      AVHOT_CUDA(cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]));
      *total_ms += ms;
      *launches += 1;
    }
  }
  return AVHOT_OK;
}

extern "C" int avhot_prof_names(char* buf, int buf_bytes) {
  AVHOT_REQUIRE(buf && buf_bytes > 0, AVHOT_ERR_INVALID, "prof_names: bad buffer");
  std::lock_guard<std::mutex> lock(avhot::g_prof_mutex);
  std::string all;
  for (auto& p : avhot::g_prof) all += p.name + "\n";
  strncpy(buf, all.c_str(), (size_t)buf_bytes - 1);
  buf[buf_bytes - 1] = 0;
  return AVHOT_OK;
}

// Diagnostic / test entry: the library's GEMM  C[M,N] = act(A[M,K] W[N,K]^T + bias)  in a given precision.
extern "C" int avhot_gemm(const float* a, const float* w, const float* bias, float* c, int M, int N, int K,
                          int relu, int precision, avhot_stream_t stream) {
  using namespace avhot;
  AVHOT_REQUIRE(precision == AVHOT_PREC_FP32 || precision == AVHOT_PREC_TC, AVHOT_ERR_UNSUPPORTED,
                "gemm: precision %d not built", precision);
  if (precision == AVHOT_PREC_FP32)
    return gemm_f32(a, nullptr, K, w, K, bias, c, N, M, N, K, relu ? ACT_RELU : ACT_NONE, as_stream(stream));
  return gemm_tf32_tc(a, K, w, K, bias, c, N, M, N, K, relu ? ACT_RELU : ACT_NONE, as_stream(stream));
}
