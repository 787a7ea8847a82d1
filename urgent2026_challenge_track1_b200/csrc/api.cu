// api.cu — error channel and device probe of the C ABI.
#include "common.cuh"
#include <stdarg.h>

namespace bsrnn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static long g_launches = 0;
void count_launches(long n) { g_launches += n; }
}  // namespace bsrnn

extern "C" const char* bsrnn_last_error(void) { return bsrnn::g_err; }
extern "C" int bsrnn_abi_version(void) { return 3; }
extern "C" int bsrnn_device_check(void) {
  // attribute queries, not cudaGetDeviceProperties: the latter costs ~3 ms per call (profiles/r02: 40 ms of a training
  // step went there) and this check guards every public entry of the host side
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
    bsrnn::set_error("no CUDA device");
    return 2;
  }
  if (major != 10) {
    bsrnn::set_error("libbsrnn_b200 is built for sm_100a only; device is sm_%d%d", major, minor);
    return 1;
  }
  return 0;
}
extern "C" long bsrnn_launch_count(int reset) {
  long v = bsrnn::g_launches;
  if (reset) bsrnn::g_launches = 0;
  return v;
}
