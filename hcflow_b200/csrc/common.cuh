// Shared helpers for the hcflow_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hcflow_b200.h"

namespace hcf {

void set_error(const char* fmt, ...);
void count_launch();
int validate_conv_args(const hcf_conv_args* a);

// Checks the launch that was just issued on `stream` (no device sync).
inline int finish_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define HCF_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::hcf::set_error(__VA_ARGS__);    \
      return HCF_EINVAL;                \
    }                                   \
  } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// logscale of every HCFlow coupling / rescaling prior: 0.318 * atan(2 s)
// (AffineCouplings.py:55,81,123,151; ConditionalFlow.py:80,89)
__device__ __forceinline__ float coupling_logscale(float s) { return 0.318f * atanf(2.0f * s); }

}  // namespace hcf
