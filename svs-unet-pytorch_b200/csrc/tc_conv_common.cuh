// Shared between the tcgen05 implicit-GEMM kernels of conv_tc.cu (one CTA per tile stream) and
// conv_tc_cluster.cu (split-K across a thread-block cluster).
#pragma once
#include "unet_internal.cuh"
#include "tc_ptx.cuh"

#include <type_traits>

namespace svs {

constexpr int kTcThreads = 192;

struct TcParams {
  const TcChunk* chunks;
  int n_chunks[4], chunk_begin[4], py[4], px[4];
  int n_phases, split_k;
  int ntw, nth;                  // M tiles along w and h of the pixel grid
  int m_tiles, n_tiles;          // tile counts (M includes the batch dimension)
  int bw, bh, nb;
  int bw_log2, bh_log2;          // log2 of bw / bh when they are powers of two, else -1
  int batch;
  int block_k;                   // K elements per chunk
  void* out;
  int out_pitch, out_coff, hout, wout, out_scale;
  const float* bias;
  int act;
  int out_flags;                 // OutFlags: accumulate into the output / keep fp32 outputs unrounded
  float* partial;
  int m_pad;                     // rows per (phase, split) slab of `partial`
  int cout;                      // GEMM N (merged deconv: 4 x channels)
  int merged;                    // 1: N = 4 phases x cout_phase channels
  int cout_phase;
  long long* dbg;                // optional per-CTA clock64 trace (8 slots per CTA), profiling only
};

// profiling timestamps: %globaltimer (ns, one clock for the whole GPU) so that CTAs on different SMs compare
__device__ __forceinline__ long long dbg_now() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Runs f(integral_constant<S>) for the runtime stage index s: inside, shared-memory descriptors are "uniform base +
// compile-time constant", so tcgen05.mma issues without per-MMA descriptor arithmetic in the vector datapath.
template <int S, int N, typename F>
__device__ __forceinline__ void dispatch_stage(int s, F&& f) {
  if constexpr (S < N) {
    if (s == S) f(std::integral_constant<int, S>{});
    else dispatch_stage<S + 1, N>(s, f);
  }
}

__device__ __forceinline__ float tc_act(float v, int act) {
  if (act == ACT_LEAKY) return v > 0.0f ? v : 0.2f * v;
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}
__device__ __forceinline__ void store16(__nv_bfloat16* dst, const float (&f)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  uint4* d = reinterpret_cast<uint4*>(dst);
  d[0] = make_uint4(w[0], w[1], w[2], w[3]);
  d[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ void store16(float* dst, const float (&f)[16]) {
  float4* d = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}


}  // namespace svs
