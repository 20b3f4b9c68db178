// C-ABI plumbing shared by all translation units: error string, version, device query.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace robir {
static thread_local char g_last_error[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
}  // namespace robir

extern "C" {

const char* robir_last_error() { return robir::g_last_error; }

int robir_abi_version() { return 1; }

// sm_count / compute capability of the current device; returns non-zero when no usable sm_100 device is present.
int robir_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  RB_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  RB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  *sm_count = prop.multiProcessorCount;
  *cc_major = prop.major;
  *cc_minor = prop.minor;
  return 0;
}

}  // extern "C"
