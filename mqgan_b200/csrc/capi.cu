// Library-level C-ABI entry points: version, thread-local error string, device check.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/mqgan_b200.h"
#include "common.cuh"

namespace mq {
static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace mq

extern "C" int mq_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* mq_last_error(void) { return mq::g_err; }

extern "C" int mq_device_check(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    mq::set_last_error("mq_device_check: no CUDA device");
    return 3;
  }
  if (prop.major != 10) {
    mq::set_last_error("mq_device_check: device %s is sm_%d%d; this library is sm_100a only",
                       prop.name, prop.major, prop.minor);
    return 3;
  }
  return 0;
}

extern "C" int mq_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}
