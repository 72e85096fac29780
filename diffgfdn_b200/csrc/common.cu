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

// Strided host -> device copy on a stream (cudaMemcpy2DAsync): the first `width_bytes` of each of `rows` rows.
// Used by the end-to-end path to move only the bins the odd-length inverse DFT reads (0..K/2, quirk Q3) out of the
// reference-layout (B, K) pinned host arrays.
extern "C" int dgfdn_copy_rows_h2d(void* dst, int64_t dst_pitch_bytes, const void* src_host, int64_t src_pitch_bytes,
                                   int64_t width_bytes, int64_t rows, void* stream) {
  DGFDN_CHECK(dst && src_host && width_bytes >= 0 && rows >= 0, "copy_rows_h2d: bad arguments");
  DGFDN_CHECK(dst_pitch_bytes >= width_bytes && src_pitch_bytes >= width_bytes, "copy_rows_h2d: pitch smaller than width");
  if (rows == 0 || width_bytes == 0) return 0;
  DGFDN_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch_bytes, src_host, (size_t)src_pitch_bytes, (size_t)width_bytes,
                               (size_t)rows, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  return 0;
}
