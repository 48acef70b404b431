// Library-level entry points of the C ABI: error slot, launch counter, device check.
#include "vt_common.cuh"

#include <cstring>

namespace vt {

static thread_local char g_err[1024] = "";
static thread_local int g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int& launch_counter() { return g_launches; }

}  // namespace vt

extern "C" {

int vt_abi_version(void) { return VT_ABI_VERSION; }
const char* vt_last_error(void) { return vt::g_err; }
int vt_last_launch_count(void) { return vt::g_launches; }

int vt_device_check(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  VT_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  VT_CUDA_OK(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (p.major != 10) {
    vt::set_error("device %s is sm_%d%d; this library is built for sm_100a only", p.name, p.major, p.minor);
    return VT_ERR_UNSUPPORTED;
  }
  return VT_OK;
}

}  // extern "C"
