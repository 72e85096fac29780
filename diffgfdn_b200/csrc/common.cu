#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dgfdn {
static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace dgfdn

extern "C" const char* dgfdn_last_error(void) { return dgfdn::g_err; }
extern "C" int dgfdn_version(void) { return 100; }
extern "C" int dgfdn_sm_count(void) { return dgfdn::sm_count(); }
