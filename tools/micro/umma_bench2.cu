// Microbenchmark 2: unrolled tcgen05.mma issue (no per-iteration scalar work) to separate issue cost from execution.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace svs;

template <int kN, int kU, int kAcc>
__global__ void __launch_bounds__(128) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    fence_proxy_async();
    constexpr uint32_t idesc = make_idesc<false, kN>();
    const uint64_t da = make_smem_desc<128>(base);
    const uint64_t db = make_smem_desc<128>(base + 32768);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (elect_one_sync()) {
#pragma unroll
        for (int u = 0; u < kU; ++u)
          umma<false>(tm + (u % kAcc) * kN, da + 2u * (u & 3), db + 2u * (u & 3), idesc, 1u);
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int kN, int kU, int kAcc> void run(int grid) {
  long long* d; cudaMalloc(&d, sizeof(long long) * grid);
  auto k = bench<kN, kU, kAcc>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 512;
  k<<<grid, 128, 100 * 1024>>>(d, iters);
  k<<<grid, 128, 100 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256]; cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d unroll=%2d accumulators=%d grid=%3d : %.1f cycles/MMA  -> %.0f MAC/clk/SM (%s)\n", kN, kU, kAcc, grid,
         double(mx) / (iters * kU), 128.0 * kN * 16 * iters * kU / double(mx), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 1, 1>(148); run<64, 4, 1>(148); run<64, 16, 1>(148); run<64, 16, 4>(148);
  run<128, 4, 1>(148); run<128, 16, 1>(148); run<128, 16, 2>(148);
  run<256, 4, 1>(148); run<256, 16, 1>(148); run<256, 16, 2>(148);
  run<32, 16, 1>(148); run<16, 16, 1>(148);
  run<256, 16, 1>(1);
  return 0;
}
