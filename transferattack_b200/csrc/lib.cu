// lib.cu — library-level entry points of libta_b200.so: version, errors, device info, launch counter,
// runtime tuning knobs. No kernels here.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "common.cuh"

namespace ta {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static std::mutex g_mu;
static std::map<std::string, int> g_tune;
static int g_sm_count[64];   // per device ordinal, 0 = not cached

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return TA_OK;
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return TA_ECUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (g_sm_count[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sm_count[dev] = n;
  }
  return g_sm_count[dev];
}

int tune_get(const char* key, int dflt) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_tune.find(key);
  return it == g_tune.end() ? dflt : it->second;
}

// cluster size for a per-sample kernel: enough CTAs to spread one sample, never more than 8 (portable limit)
int pick_cluster(int64_t n, int threads) {
  const int forced = tune_get("reduce.cluster", 0);
  if (forced > 0) return forced;
  const int64_t per_cta = (int64_t)threads * 4 * 4;   // >= 4 vectors per thread before splitting further
  int cl = 1;
  while (cl < 8 && n / (cl * 2) >= per_cta) cl *= 2;
  return cl;
}

}  // namespace ta

extern "C" {

int ta_version(void) { return TA_ABI_VERSION; }

const char* ta_last_error(void) { return ta::g_err; }

int ta_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  cudaDeviceProp p;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) {
    ta::set_error("ta_device_info: CUDA error %d (%s)", (int)e, cudaGetErrorString(e));
    cudaGetLastError();
    return TA_ECUDA;
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return TA_OK;
}

int64_t ta_launch_count(void) { return ta::g_launches.load(std::memory_order_relaxed); }

// Not part of the reference-facing surface: runtime knobs used by the benchmark sweep
// (e.g. "fused.cluster", "fused.threads", "fused.variant"). Unknown keys are stored and ignored.
int ta_tune_set(const char* key, int value) {
  if (!key) return TA_EINVAL;
  std::lock_guard<std::mutex> lk(ta::g_mu);
  ta::g_tune[key] = value;
  return TA_OK;
}

}  // extern "C"
