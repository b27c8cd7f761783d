#include "common.cuh"
#include <stdarg.h>
#include "../../include/dt_b200.h"

namespace dt {
static thread_local char g_err[512] = "";
char* last_error_buf() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace dt

extern "C" const char* dt_last_error(void) { return dt::last_error_buf(); }
extern "C" int dt_abi_version(void) { return DT_B200_ABI_VERSION; }

// Asynchronous fill of a caller-owned device buffer on the caller's stream (a memset node when the stream is being
// captured into a CUDA graph): the hot path zeroes its count / output buffers with this instead of library kernels.
extern "C" int dt_memset(void* ptr, int value, size_t bytes, void* stream) {
  if (bytes == 0) return 0;
  DT_CHECK_ARG(ptr != nullptr, "dt_memset: null pointer");
  DT_CHECK_CUDA(cudaMemsetAsync(ptr, value, bytes, (cudaStream_t)stream));
  return 0;
}
