// Shared helpers for the DetectAndTrack B200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace dt {

// Last error text, one per host thread (dt_last_error()).
char* last_error_buf();
void set_error(const char* fmt, ...);

#define DT_CHECK_ARG(cond, ...)                         \
  do {                                                  \
    if (!(cond)) {                                      \
      ::dt::set_error(__VA_ARGS__);                     \
      return 1;                                         \
    }                                                   \
  } while (0)

#define DT_CHECK_CUDA(expr)                                                    \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      ::dt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                      __FILE__, __LINE__);                                     \
      return 2;                                                                \
    }                                                                          \
  } while (0)

#define DT_CHECK_LAUNCH() DT_CHECK_CUDA(cudaGetLastError())

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace dt
