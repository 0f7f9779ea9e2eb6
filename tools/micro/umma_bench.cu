// Microbenchmark: tcgen05.mma issue/execute rate (cta_group::1, M=128, bf16) from fixed smem operands.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I svs-unet-pytorch_b200/csrc tools/micro/umma_bench.cu -o gpurun_out/umma_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace svs;

template <int kN>
__global__ void __launch_bounds__(128) bench(long long* out, int n_mma, int commit_every, uint32_t a_shift, uint32_t sbo, int rot) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    fence_proxy_async();
    constexpr uint32_t idesc = make_idesc<false, kN>();
    const uint32_t a_addr = base + a_shift;
    const uint64_t da = static_cast<uint64_t>((a_addr & 0x3FFFF) >> 4) | (1ull << 16) | (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint64_t db = make_smem_desc<128>(base + 32768);
    long long t0 = clock64();
    uint32_t par = 0;
    int since = 0;
    for (int i = 0; i < n_mma; ++i) {
      if (elect_one_sync()) umma<false>(tm + (i % rot) * kN, da + 2u * ((i / rot) & 3), db + 2u * ((i / rot) & 3), idesc, i >= rot ? 1u : 0u);
      if (++since == commit_every) { since = 0; if (elect_one_sync()) umma_commit(smem_u32(&bar)); __syncwarp(); mbar_wait(smem_u32(&bar), par); par ^= 1; }
    }
    if (elect_one_sync()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), par);
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int kN> void run(const char* name, int commit_every, uint32_t a_shift, uint32_t sbo, int grid, int rot = 1) {
  long long* d; cudaMalloc(&d, sizeof(long long) * grid);
  auto k = bench<kN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int n = 4096;
  k<<<grid, 128, 100 * 1024>>>(d, n, commit_every, a_shift, sbo, rot);
  k<<<grid, 128, 100 * 1024>>>(d, n, commit_every, a_shift, sbo, rot);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256]; cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-28s N=%3d rot=%d commit_every=%4d shift=%5u sbo=%5u grid=%3d : %.1f cycles/MMA (%s)\n", name, kN, rot, commit_every, a_shift, sbo, grid,
         double(mx) / n, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64>("aligned 1 CTA", 4096, 0, 1024, 1);
  run<128>("aligned 1 CTA", 4096, 0, 1024, 1);
  run<256>("aligned 1 CTA", 4096, 0, 1024, 1);
  run<64>("aligned all SMs", 4096, 0, 1024, 148);
  run<128>("aligned all SMs", 4096, 0, 1024, 148);
  run<256>("aligned all SMs", 4096, 0, 1024, 148);
  run<64>("shifted+sbo1280 all SMs", 4096, 128 * 11, 1280, 148);
  run<128>("shifted+sbo1280 all SMs", 4096, 128 * 11, 1280, 148);
  run<256>("shifted+sbo1280 all SMs", 4096, 128 * 11, 1280, 148);
  run<128>("commit every 4", 4, 0, 1024, 148);
  run<128>("commit every 36", 36, 0, 1024, 148);
  run<64>("commit every 36", 36, 0, 1024, 148);
  run<64>("rotate 2 accumulators", 4096, 0, 1024, 148, 2);
  run<64>("rotate 4 accumulators", 4096, 0, 1024, 148, 4);
  run<64>("rotate 8 accumulators", 4096, 0, 1024, 148, 8);
  run<128>("rotate 2 accumulators", 4096, 0, 1024, 148, 2);
  run<128>("rotate 4 accumulators", 4096, 0, 1024, 148, 4);
  run<256>("rotate 2 accumulators", 4096, 0, 1024, 148, 2);
  run<32>("rotate 8 accumulators", 4096, 0, 1024, 148, 8);
  return 0;
}
