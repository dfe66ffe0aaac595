// Error reporting for the C ABI: every export returns 0 on success, <0 for a bad argument,
// >0 for a cudaError_t; b4cp_last_error() returns the thread-local message.
#include <atomic>
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

static thread_local char g_err[512] = "";
static std::atomic<long> g_launches{0};

void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_last_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return (int)e;
}

}  // namespace b4cp

extern "C" const char* b4cp_last_error(void) { return b4cp::g_err; }

extern "C" long b4cp_launch_count(void) { return b4cp::g_launches.load(); }

extern "C" int b4cp_version(void) { return B4CP_VERSION; }

extern "C" int b4cp_device_check(void) {
  int dev = 0;
  B4CP_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  B4CP_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    b4cp::set_last_error("libb4cp is built for sm_100a only; device %d is sm_%d%d", dev,
                         prop.major, prop.minor);
    return -10;
  }
  return 0;
}
