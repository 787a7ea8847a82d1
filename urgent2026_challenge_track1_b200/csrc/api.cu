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
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    bsrnn::set_error("no CUDA device");
    return 2;
  }
  if (p.major != 10) {
    bsrnn::set_error("libbsrnn_b200 is built for sm_100a only; device is sm_%d%d", p.major, p.minor);
    return 1;
  }
  return 0;
}
extern "C" long bsrnn_launch_count(int reset) {
  long v = bsrnn::g_launches;
  if (reset) bsrnn::g_launches = 0;
  return v;
}
