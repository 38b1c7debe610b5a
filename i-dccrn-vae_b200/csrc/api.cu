// Error reporting and device queries of the C ABI (include/idv.h).
#include <stdarg.h>

#include "idv_common.cuh"

namespace idv {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace idv

extern "C" int idv_abi_version(void) { return IDV_ABI_VERSION; }
extern "C" const char* idv_last_error(void) { return idv::g_err; }
extern "C" int idv_device_sm_count(int* out) {
  using namespace idv;
  IDV_CHECK_ARG(out, "idv_device_sm_count: null pointer");
  int dev = 0;
  IDV_CUDA(cudaGetDevice(&dev));
  IDV_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return IDV_OK;
}
