// Shared internals of libsvs_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "svs_b200.h"

namespace svs {

// ---- thread-local error reporting -----------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define SVS_CUDA_TRY(expr)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      return ::svs::fail(SVS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
    }                                                                                          \
  } while (0)

#define SVS_CHECK_LAUNCH(name)                                                                 \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      return ::svs::fail(SVS_ERR_CUDA, std::string(name) + " launch: " + cudaGetErrorString(_e)); \
    }                                                                                          \
  } while (0)

#define SVS_REQUIRE(cond, msg)                                                                 \
  do {                                                                                         \
    if (!(cond)) return ::svs::fail(SVS_ERR_INVALID_ARG, std::string(msg));                   \
  } while (0)

// ---- per-device constant tables (twiddles, Hann window, OLA envelope) -----------------------
struct SpectralTables {
  const float2* tw1024;     // W_1024^m = exp(-2 pi i m / 1024), m in [0,1024)
  const float* hann;        // periodic Hann, 1024 (float64 rounded to float32)
  const float* env_both;    // [256]  1 / fl32(fl32(w^2[r+768]) + w^2[r])   two frames cover the sample
  const float* env_single;  // [768]  1 / fl32(w^2[r])                       one frame covers the sample
};
int get_spectral_tables(SpectralTables* out);   // lazily builds the tables on the current device

int num_sms();

// ---- small device helpers --------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// max of non-negative floats through the integer ordering (deterministic: max is order independent)
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}

}  // namespace svs
