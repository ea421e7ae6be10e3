// Process-level plumbing of libhmz.so: error strings, launch counter, device queries.
#include "hmz_common.cuh"

namespace hmz {

std::atomic<long long> g_launches{0};

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = kSmFallback;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return kSmFallback;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace hmz

extern "C" {

const char* hmz_last_error(void) { return hmz::error_buffer(); }

int hmz_version(void) { return 100; }

int64_t hmz_launch_count(void) { return (int64_t)hmz::g_launches.load(); }

int hmz_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return hmz::fail(HMZ_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return hmz::fail(HMZ_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return HMZ_OK;
}

}  // extern "C"
