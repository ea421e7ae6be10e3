// Process-level plumbing of libhmz.so: error strings, launch counter, device queries.
#include <mutex>
#include <vector>

#include "hmz_common.cuh"

namespace hmz {

std::atomic<int> g_prof_on{0};
namespace {
struct ProfRec {
  int cls;
  cudaEvent_t start, stop;
};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
}  // namespace

void prof_push(int cls, cudaStream_t stream, bool start) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (start) {
    ProfRec r{cls, nullptr, nullptr};
    if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
    cudaEventRecord(r.start, stream);
    g_prof.push_back(r);
  } else {
    // the matching record is the most recent one of this class (scopes of one class do not nest)
    for (size_t i = g_prof.size(); i-- > 0;)
      if (g_prof[i].cls == cls) {
        cudaEventRecord(g_prof[i].stop, stream);
        break;
      }
  }
}

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

int hmz_version(void) { return 200; }

int64_t hmz_launch_count(void) { return (int64_t)hmz::g_launches.load(); }

int hmz_prof_begin(void) {
  std::lock_guard<std::mutex> lk(hmz::g_prof_mu);
  for (auto& r : hmz::g_prof) {
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  hmz::g_prof.clear();
  hmz::g_prof_on.store(1);
  return HMZ_OK;
}

int hmz_prof_end(double* ms_by_class, int64_t* launches_by_class) {
  hmz::g_prof_on.store(0);
  std::lock_guard<std::mutex> lk(hmz::g_prof_mu);
  for (int c = 0; c < HMZ_PROF_CLASSES; ++c) {
    if (ms_by_class) ms_by_class[c] = 0.0;
    if (launches_by_class) launches_by_class[c] = 0;
  }
  int rc = HMZ_OK;
  for (auto& r : hmz::g_prof) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(r.stop);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.start, r.stop);
    if (e != cudaSuccess) rc = hmz::fail(HMZ_ERR_CUDA, "hmz_prof_end: %s", cudaGetErrorString(e));
    if (r.cls >= 0 && r.cls < HMZ_PROF_CLASSES) {
      if (ms_by_class) ms_by_class[r.cls] += ms;
      if (launches_by_class) launches_by_class[r.cls] += 1;
    }
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  hmz::g_prof.clear();
  return rc;
}

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
