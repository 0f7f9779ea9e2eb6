// tc_conv_ck_kernel: the TMA-im2col implicit GEMM of conv_tc.cu for the DEEP layers (conv5, conv6, deconv1) with
// split-K ACROSS A THREAD-BLOCK CLUSTER.
//
// The deep layers have too few output tiles to fill 148 SMs, so K is split; conv_tc.cu writes fp32 partial tiles
// to global memory and a second kernel reduces them.  Here the kSplit CTAs of a cluster own the K slices of ONE
// 128 x 128 output tile: each CTA parks its fp32 accumulator in its own shared memory (the drained operand
// ring), and after one cluster barrier finishes 128 / kSplit rows of the tile by reading the kSplit slices over
// distributed shared memory in fixed rank order (deterministic), then bias + activation and coalesced stores (a
// warp writes one pixel's 128 channels).  No partial buffer, no second launch.
//
// Measured alternatives:
//  * PUSHING rows into the owner (st.async / st.shared::cluster during the TMEM drain) instead of pulling them:
//    slower -- with every SM pushing at once remote stores moved ~7 B/clk/SM, the pulls below ~2x that;
//  * 2 x 2 clusters sharing operand tiles by TMA multicast: slower than one CTA per tile (+1..2 us per layer), the
//    CTAs of a cluster advance in lock step and L2 already de-duplicates concurrent requests for the same lines;
//  * the reduction loop runs with ONE warp per scheduler and nothing to overlap with, so it is bound by
//    dependent-issue latency: branches in the activation and integer divisions in the pixel decode cost more than
//    the DSMEM traffic (3.8 us -> see DESIGN.md after making it branch-free).
#include "tc_conv_common.cuh"


namespace svs {


__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}


// N tile 128 (six 32 KB stages) or 256 (four 48 KB stages): the deep layers are bound by bytes INTO shared memory,
// and a 128 x 256 tile moves 48 KB per K chunk for twice the MACs of a 128 x 128 tile's 32 KB.  With the tile count
// halved the K split doubles, so the grid still fills the SMs (conv5: 32 tiles x 4, conv6: 16 x 8, deconv1: 32 x 4).

constexpr size_t kCkSmemBytes = static_cast<size_t>(6) * (128 + 128) * 128 + 1024 + 256 + 2048;   // same for both shapes
static_assert(static_cast<size_t>(4) * (128 + 256) * 128 == static_cast<size_t>(6) * (128 + 128) * 128, "ring bytes");


__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// not volatile: the compiler may interleave these with the arithmetic of earlier rows.  `addr` is derived from
// values pinned behind the cluster barrier (see the asm fence after it), so they cannot be hoisted above it.
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
  float4 v;
  asm("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_smem_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// accumulator parking lot: row r = 4 * kBlockN bytes, 16-byte chunk c at (c ^ (r & 31)): conflict-free both for the
// row-per-thread TMEM drain and for the chunk-per-thread reduction (the XOR stays inside a 32-chunk half)
template <int kBlockN>
__device__ __forceinline__ uint32_t park_off(int r, int c) { return static_cast<uint32_t>(r * (4 * kBlockN) + ((c ^ (r & 31)) << 4)); }

__device__ __forceinline__ void store4(__nv_bfloat16* dst, const float (&o)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
  __nv_bfloat162 c = __floats2bfloat162_rn(o[2], o[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&c);
  *reinterpret_cast<uint2*>(dst) = u;
}
__device__ __forceinline__ void store4(float* dst, const float (&o)[4], int out_flags = 0) {
  float4 r = make_float4(o[0], o[1], o[2], o[3]);
  if (out_flags & OUT_ACCUMULATE) {
    const float4 old = *reinterpret_cast<const float4*>(dst);
    r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
  }
  if (!(out_flags & OUT_KEEP_FP32)) r = make_float4(round_tf32(r.x), round_tf32(r.y), round_tf32(r.z), round_tf32(r.w));
  *reinterpret_cast<float4*>(dst) = r;
}
__device__ __forceinline__ void store4(__nv_bfloat16* dst, const float (&o)[4], int) { store4(dst, o); }

template <typename OutT, bool kTf32, int kSplit, int kBlockN>
__global__ void __launch_bounds__(kTcThreads)
tc_conv_ck_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b0,
                  const __grid_constant__ CUtensorMap tmap_b1, const __grid_constant__ CUtensorMap tmap_b2,
                  const __grid_constant__ CUtensorMap tmap_b3, const TcParams p) {
  constexpr int kStages = kBlockN == 128 ? 6 : 4, kSwz = 128;
  constexpr int kABytes = 128 * kSwz, kBBytes = kBlockN * kSwz, kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = kBlockN;
  constexpr int kRowsPerCta = 128 / kSplit;
  static_assert(kStages * kStageBytes >= 128 * 4 * kBlockN, "the operand ring doubles as the accumulator parking lot");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * kStageBytes + 8 * (2 * kStages + 4));
  float* sbias = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes + 256);
  for (int i = threadIdx.x; i < p.cout; i += kTcThreads) sbias[i] = __ldg(&p.bias[i]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();            // == K slice
  const int tile = blockIdx.x / kSplit;            // (phase, n tile, m tile)
  const int per_z = p.m_tiles * p.n_tiles;
  const int phase = tile / per_z, rem = tile - phase * per_z;
  const int nt = rem / p.m_tiles, mt = rem - nt * p.m_tiles;
  const int n_chunks = p.n_chunks[phase];
  const int per_split = (n_chunks + kSplit - 1) / kSplit;
  const int c_begin = static_cast<int>(rank) * per_split;
  const int n_iter = max(0, min(n_chunks, c_begin + per_split) - c_begin);
  const int tw = mt % p.ntw, th = (mt / p.ntw) % p.nth, tb = mt / (p.ntw * p.nth);

  long long* dbg = p.dbg ? p.dbg + 8 * blockIdx.x : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = dbg_now();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b0);
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  if (dbg && threadIdx.x == 0) dbg[1] = dbg_now();
  pdl_wait();

  if (warp == 0) {
    const CUtensorMap* tb_map = phase == 0 ? &tmap_b0 : phase == 1 ? &tmap_b1 : phase == 2 ? &tmap_b2 : &tmap_b3;
    const TcChunk* chunks = p.chunks + p.chunk_begin[phase];
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kStages;
      mbar_wait(empty_bar(s), ((it / kStages) & 1) ^ 1);
      const int ci = c_begin + it;
      const TcChunk ch = chunks[ci];
      const uint32_t a_dst = smem_base + s * kStageBytes;
      if (elect_one_sync()) {
        mbar_expect_tx(full_bar(s), kStageBytes);
        tma_load_5d(a_dst, &tmap_a, full_bar(s), ch.c_inner, tw * p.bw + ch.dw, ch.ph, th * p.bh + ch.dh, tb * p.nb);
        tma_load_2d(a_dst + kABytes, tb_map, full_bar(s), ci * p.block_k, nt * kBlockN);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc<kTf32, kBlockN>();
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kStages;
      mbar_wait(full_bar(s), (it / kStages) & 1);
      if (dbg && it == 0 && lane == 0) dbg[2] = dbg_now();
      tc_fence_after();
      const uint32_t acc0 = it > 0 ? 1u : 0u;
      dispatch_stage<0, kStages>(s, [&](auto sc) {
        constexpr int S = decltype(sc)::value;
        const uint32_t a_addr = smem_base + S * kStageBytes;
        const uint64_t da = make_smem_desc<kSwz>(a_addr);
        const uint64_t db = make_smem_desc<kSwz>(a_addr + kABytes);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<kTf32>(tmem_base, da + 2u * k, db + 2u * k, idesc, k > 0 ? 1u : acc0);
          umma_commit(bar_base + 8u * (kStages + S));
        }
        __syncwarp();
      });
    }
    if (elect_one_sync()) umma_commit(tmem_full_bar);
    __syncwarp();
    if (dbg && lane == 0) dbg[3] = dbg_now();
  } else {
    // drain: thread = accumulator row.  Every MMA has retired when tmem_full fires, hence every TMA write
    // it consumed has landed and the operand ring is free to hold the 64 KB fp32 tile.
    const int q = warp & 3;
    const int r = 32 * q + lane;
    mbar_wait(tmem_full_bar, 0);
    if (dbg && threadIdx.x == 64) dbg[4] = dbg_now();
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16);
#pragma unroll 2
    for (int c = 0; c < kBlockN; c += 16) {
      uint32_t v[16];
      if (n_iter > 0) {
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        st_smem_f4(smem_base + park_off<kBlockN>(r, (c >> 2) + i), __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
    }
    tc_fence_before();
  }
  __syncwarp();
  cluster_sync_all();                              // every K slice of the tile is parked (release / acquire)
  if (dbg && threadIdx.x == 64) dbg[7] = dbg_now();
  uint32_t smem_after = smem_base;                 // addresses used below are "produced" after the barrier
  asm volatile("" : "+r"(smem_after)::"memory");

  if (warp >= 2) {
    // reduce rows [rank * kRowsPerCta, +kRowsPerCta): a warp handles one row per step, lane = 4-channel chunk
    const int e = threadIdx.x - 64;
    const int chunk = e & 31;
    OutT* const out_base = reinterpret_cast<OutT*>(p.out);
    const int py = p.py[phase], px = p.px[phase];
    // One warp per scheduler and nothing left to overlap with: the loop is bound by dependent-issue latency, so
    // it is kept short (branch-free activation, shifts for the pixel decode) and kBatch rows are interleaved.
    const float slope = p.act == ACT_LEAKY ? 0.2f : (p.act == ACT_RELU ? 0.0f : 1.0f);
    uint32_t peer[kSplit];
#pragma unroll
    for (int k = 0; k < kSplit; ++k) peer[k] = map_to_cta(smem_after, k);
    const int sh = p.bw_log2 + p.bh_log2;
    const size_t pix_pitch = static_cast<size_t>(p.out_pitch);
    constexpr int kBatch = 16 / kSplit;              // rows in flight: 16 loads per thread
    static_assert((kRowsPerCta / 4) % kBatch == 0, "rows per CTA must be a multiple of the load batch");
#pragma unroll 1
    for (int half = 0; half < kBlockN / 128; ++half) {     // a warp covers 128 channels of a row per step
      const int n = nt * kBlockN + 128 * half + 4 * chunk;
      const float4 bv = *reinterpret_cast<const float4*>(&sbias[n]);
      OutT* const out_n = out_base + p.out_coff + n;
#pragma unroll 1
      for (int i0 = 0; i0 < kRowsPerCta / 4; i0 += kBatch) {
        float4 v[kBatch][kSplit];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const int r = static_cast<int>(rank) * kRowsPerCta + 4 * (i0 + j) + (e >> 5);
          const uint32_t off = park_off<kBlockN>(r, 32 * half + chunk);
#pragma unroll
          for (int k = 0; k < kSplit; ++k) v[j][k] = ld_cluster_f4(peer[k] + off);
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const int r = static_cast<int>(rank) * kRowsPerCta + 4 * (i0 + j) + (e >> 5);
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < kSplit; ++k) {          // fixed order: deterministic
            acc.x += v[j][k].x; acc.y += v[j][k].y; acc.z += v[j][k].z; acc.w += v[j][k].w;
          }
          const int iw = r & (p.bw - 1), ih = (r >> p.bw_log2) & (p.bh - 1), ib = r >> sh;
          const int gx = tw * p.bw + iw, gy = th * p.bh + ih, b = tb * p.nb + ib;
          const int oy = gy * p.out_scale + py, ox = gx * p.out_scale + px;
          // max(v, slope * v + 0): LeakyReLU (0.2), ReLU (0) and identity (1) without a branch
          const float t[4] = {acc.x + bv.x, acc.y + bv.y, acc.z + bv.z, acc.w + bv.w};
          const float o[4] = {fmaxf(t[0], fmaf(slope, t[0], 0.0f)), fmaxf(t[1], fmaf(slope, t[1], 0.0f)),
                              fmaxf(t[2], fmaf(slope, t[2], 0.0f)), fmaxf(t[3], fmaf(slope, t[3], 0.0f))};
          if (b < p.batch) store4(out_n + static_cast<size_t>((b * p.hout + oy) * p.wout + ox) * pix_pitch, o, p.out_flags);
        }
      }
    }
  }
  __syncwarp();
  if (dbg && threadIdx.x == 64) dbg[5] = dbg_now();
  // peers have finished reading this CTA's parking lot: their loads have returned, so a relaxed arrive is enough
  // (a releasing arrive is a GPU-scope fence that would wait for the output stores above)
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  if (dbg && threadIdx.x == 0) dbg[6] = dbg_now();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

bool ck_supported(const TcLayer& t, int split, int block_n) {
  const bool pow2 = !(t.bw & (t.bw - 1)) && !(t.bh & (t.bh - 1));     // pixel decode by shifts in the reduction
  const bool shape = block_n == 128 ? t.block_n == 128 : (block_n == 256 && t.has_wide);
  return t.enabled && !t.merged && pow2 && t.swz == 128 && shape && (split == 2 || split == 4 || split == 8);
}

template <typename OutT, bool kTf32, int kSplit, int kBlockN>
static int launch_ck(const CUtensorMap& ta, const TcLayer& t, const TcParams& p, cudaStream_t st) {
  auto kern = tc_conv_ck_kernel<OutT, kTf32, kSplit, kBlockN>;
  SVS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kCkSmemBytes)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.m_tiles * p.n_tiles * p.n_phases * kSplit);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = kCkSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kSplit; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  const CUtensorMap* tb = kBlockN == 256 ? t.tmap_b_wide : t.tmap_b;
  SVS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, tb[0], tb[1], tb[2], tb[3], p));
  return SVS_OK;
}

int ck_launch_layer(const TcLayer& t, bool tf32, int block_n, const CUtensorMap& ta, const TcParams& p, cudaStream_t st) {
#define SVS_CK_CASE(S, N)                                                                  \
  if (p.split_k == S && block_n == N)                                                      \
    return tf32 ? launch_ck<float, true, S, N>(ta, t, p, st) : launch_ck<__nv_bfloat16, false, S, N>(ta, t, p, st);
  SVS_CK_CASE(2, 128)
  SVS_CK_CASE(4, 128)
  SVS_CK_CASE(8, 128)
  SVS_CK_CASE(2, 256)
  SVS_CK_CASE(4, 256)
  SVS_CK_CASE(8, 256)
#undef SVS_CK_CASE
  return fail(SVS_ERR_NOT_IMPLEMENTED, "ck_launch_layer: unsupported split / tile");
}

}  // namespace svs
