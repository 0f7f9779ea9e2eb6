// Shared internals of libsvs_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>

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

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// Every layer kernel is launched with programmaticStreamSerialization so that its prologue (barrier
// init, TMEM allocation, tensor-map prefetch, resident weights) overlaps the tail of the previous
// layer; inside the kernel pdl_wait() blocks until the previous grid has completed and flushed, and
// must precede the first access to anything that grid wrote (or reads and this grid overwrites).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
