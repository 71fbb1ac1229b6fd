// Version / error reporting entry points of the C ABI (include/imp_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace imp {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace imp

extern "C" int imp_version(void) { return IMP_VERSION; }
extern "C" const char* imp_last_error_string(void) { return imp::g_err; }
extern "C" int imp_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}
