// Error reporting and device queries for libtripsb200.
#include "tb200_common.cuh"
#include <cstdarg>
#include <cstring>

namespace tb200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace tb200

extern "C" {

const char* tb200_last_error(void) { return tb200::g_err; }

int tb200_version(void) { return 100; }

// Fills sm_count, l2_bytes, total_mem_bytes, cc (major*10+minor) for the current device.
int tb200_device_info(int* sm_count, int64_t* l2_bytes, int64_t* mem_bytes, int* cc) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    tb200::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return (int)e;
  }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) {
    tb200::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return (int)e;
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (l2_bytes) *l2_bytes = p.l2CacheSize;
  if (mem_bytes) *mem_bytes = (int64_t)p.totalGlobalMem;
  if (cc) *cc = p.major * 10 + p.minor;
  return 0;
}

// The library carries sm_100a SASS only; refuse anything else loudly instead of failing at first launch.
int tb200_require_sm100(void) {
  int cc = 0;
  int rc = tb200_device_info(nullptr, nullptr, nullptr, &cc);
  if (rc) return rc;
  if (cc / 10 != 10) {
    tb200::set_error("libtripsb200 is built for sm_100a only; device reports compute capability %d.%d", cc / 10, cc % 10);
    return TB200_ENOTSM100;
  }
  return 0;
}

}  // extern "C"
