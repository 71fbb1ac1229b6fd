// Shared helpers of the imp_b200 library (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "imp_b200.h"

namespace imp {

void set_error(const char* fmt, ...);

#define IMP_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::imp::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

#define IMP_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::imp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

#define IMP_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      ::imp::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace imp
