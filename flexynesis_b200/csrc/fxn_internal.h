// Internal helpers shared by the .cu translation units of libfxn_b200.so
#pragma once
#include "../../include/flexynesis_b200.h"
#include <cuda_runtime.h>

namespace fxn {
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define FXN_CHECK_LAUNCH(what)                                                          \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return set_error(FXN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e__)); \
    count_launch();                                                                     \
  } while (0)

inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }
}  // namespace fxn
