// Thread-block-cluster variants of the TMA-im2col implicit GEMM of conv_tc.cu for the DEEP layers (conv5, conv6,
// deconv1, deconv2).
//
// (1) tc_conv_ck_kernel: split-K ACROSS THE CLUSTER.  The deep layers have too few output tiles to fill 148
// SMs, so K is split; conv_tc.cu writes fp32 partial tiles to global memory and a second kernel reduces them.
// Here the kSplit CTAs of a cluster own the K slices of ONE output tile, park their fp32 accumulators in their
// own shared memory (the drained operand ring), and after a cluster barrier each CTA reduces 128/kSplit rows
// over distributed shared memory in fixed rank order (deterministic), applies bias + activation and stores.  No
// partial buffer, no second launch, and the stores are fully coalesced (a warp writes one pixel's 128 channels).
//
// (2) tc_conv_mc_kernel (opt-in, SVS_TC_CLUSTER=1): operand tiles shared across a cluster by TMA multicast.
//
// Those layers have few pixels and large weights: 128 CTAs each stream a [128 px x 64 ch] A chunk and a
// [128 co x 64 ch] B chunk per step and run at ~64 B/clk/SM, the L2 -> SMEM limit, i.e. half of what the
// 128 x 128 MMA tile consumes.  In a 2 x 2 cluster (2 adjacent M tiles x 2 adjacent N tiles, same K range)
// CTA (mi, ni) loads only HALF of its A tile (rows 64 ni ..) and multicasts it to the two CTAs of its M tile,
// and half of its B tile (rows 64 mi ..) multicast to the two CTAs of its N tile: every CTA still receives a
// full 32 KB stage but issues 16 KB of L2 reads.  With one N tile (deconv2) the cluster is 2 x 1 and only B
// is shared (24 KB).
//
// Barrier protocol per stage: full[s] (count 1 + 32 KB of tx from up to three CTAs); empty[s] counts one
// tcgen05.commit arrival from every CTA that writes into this CTA's stage (itself and its partners), issued
// as a multicast commit, so a producer never overwrites a partner's operands that are still being read.
#include "tc_conv_common.cuh"

#include <mutex>

namespace svs {

extern long long* g_tc_dbg;
extern int g_tc_dbg_layer;
void tc_tiling(const TcLayer& t, const LayerGeom& g, int batch, int* m_tiles, int* split_k);

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                               int c3, int c4, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

template <int S, int N, typename F>
__device__ __forceinline__ void dispatch_stage_mc(int s, F&& f) {
  if constexpr (S < N) {
    if (s == S) f(std::integral_constant<int, S>{});
    else dispatch_stage_mc<S + 1, N>(s, f);
  }
}

constexpr int kMcBlockN = 128;
constexpr int kMcStages = 6;
constexpr size_t kMcSmemBytes = static_cast<size_t>(kMcStages) * (128 + kMcBlockN) * 128 + 1024 + 256 + 2048;


// =================================================================================================
// (1) split-K across the cluster
// =================================================================================================
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_smem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_smem_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// accumulator parking lot: row r = 512 bytes, 16-byte chunk c at (c ^ (r & 31)): conflict-free both for the
// row-per-thread TMEM drain and for the chunk-per-thread reduction
__device__ __forceinline__ uint32_t park_off(int r, int c) { return static_cast<uint32_t>(r * 512 + ((c ^ (r & 31)) << 4)); }

__device__ __forceinline__ void store4(__nv_bfloat16* dst, const float (&o)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
  __nv_bfloat162 c = __floats2bfloat162_rn(o[2], o[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&c);
  *reinterpret_cast<uint2*>(dst) = u;
}
__device__ __forceinline__ void store4(float* dst, const float (&o)[4]) {
  *reinterpret_cast<float4*>(dst) = make_float4(round_tf32(o[0]), round_tf32(o[1]), round_tf32(o[2]), round_tf32(o[3]));
}

template <typename OutT, bool kTf32, int kSplit>
__global__ void __launch_bounds__(kTcThreads)
tc_conv_ck_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b0,
                  const __grid_constant__ CUtensorMap tmap_b1, const __grid_constant__ CUtensorMap tmap_b2,
                  const __grid_constant__ CUtensorMap tmap_b3, const TcParams p) {
  constexpr int kBlockN = kMcBlockN, kStages = kMcStages, kSwz = 128;
  constexpr int kABytes = 128 * kSwz, kBBytes = kBlockN * kSwz, kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = kBlockN;
  constexpr int kRowsPerCta = 128 / kSplit;
  static_assert(kStages * kStageBytes >= 128 * 512, "the operand ring doubles as the accumulator parking lot");

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
      dispatch_stage_mc<0, kStages>(s, [&](auto sc) {
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
        st_smem_f4(smem_base + park_off(r, (c >> 2) + i), __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
    }
    tc_fence_before();
  }
  __syncwarp();
  cluster_sync_all();                              // every K slice of the tile is parked (release / acquire)
  if (dbg && threadIdx.x == 64) dbg[7] = dbg_now();

  if (warp >= 2) {
    // reduce rows [rank * kRowsPerCta, +kRowsPerCta): a warp handles one row per step, lane = 4-channel chunk
    const int e = threadIdx.x - 64;
    const int chunk = e & 31;
    const int n = nt * kBlockN + 4 * chunk;
    const float4 bv = *reinterpret_cast<const float4*>(&sbias[n]);
    OutT* const out_base = reinterpret_cast<OutT*>(p.out);
    const int py = p.py[phase], px = p.px[phase];
    uint32_t peer[kSplit];
#pragma unroll
    for (int k = 0; k < kSplit; ++k) peer[k] = map_to_cta(smem_base, k);
    // kBatch rows x kSplit slices = 16 DSMEM loads in flight per thread; the loads of a batch are issued before
    // any of its stores (volatile asm keeps program order, and a remote load costs several hundred ns)
    constexpr int kBatch = 16 / kSplit;
    static_assert((kRowsPerCta / 4) % kBatch == 0, "rows per CTA must be a multiple of the load batch");
#pragma unroll 1
    for (int i0 = 0; i0 < kRowsPerCta / 4; i0 += kBatch) {
      float4 v[kBatch][kSplit];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int r = static_cast<int>(rank) * kRowsPerCta + 4 * (i0 + j) + (e >> 5);
        const uint32_t off = park_off(r, chunk);
#pragma unroll
        for (int k = 0; k < kSplit; ++k)            // own slice through the local port: DSMEM moves ~20 B/clk/SM
          v[j][k] = (k == static_cast<int>(rank)) ? ld_smem_f4(smem_base + off) : ld_dsmem_f4(peer[k] + off);
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int r = static_cast<int>(rank) * kRowsPerCta + 4 * (i0 + j) + (e >> 5);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < kSplit; ++k) {          // fixed order: deterministic
          acc.x += v[j][k].x; acc.y += v[j][k].y; acc.z += v[j][k].z; acc.w += v[j][k].w;
        }
        const int iw = r & (p.bw - 1), ih = (r >> p.bw_log2) & (p.bh - 1), ib = r >> (p.bw_log2 + p.bh_log2);
        const int gx = tw * p.bw + iw, gy = th * p.bh + ih, b = tb * p.nb + ib;
        const int oy = gy * p.out_scale + py, ox = gx * p.out_scale + px;
        const float o[4] = {tc_act(acc.x + bv.x, p.act), tc_act(acc.y + bv.y, p.act), tc_act(acc.z + bv.z, p.act),
                            tc_act(acc.w + bv.w, p.act)};
        if (b < p.batch)
          store4(out_base + ((static_cast<size_t>(b) * p.hout + oy) * p.wout + ox) * p.out_pitch + p.out_coff + n, o);
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

bool ck_supported(const svs_unet_plan* plan, int li, int split) {
  const TcLayer& t = plan->tc[li];
  const bool pow2 = !(t.bw & (t.bw - 1)) && !(t.bh & (t.bh - 1));     // pixel decode by shifts in the reduction
  return t.enabled && !t.merged && pow2 && t.swz == 128 && t.block_n == 128 && (split == 2 || split == 4 || split == 8);
}

template <typename OutT, bool kTf32, int kSplit>
static int launch_ck(const CUtensorMap& ta, const TcLayer& t, const TcParams& p, cudaStream_t st) {
  auto kern = tc_conv_ck_kernel<OutT, kTf32, kSplit>;
  SVS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMcSmemBytes)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.m_tiles * p.n_tiles * p.n_phases * kSplit);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = kMcSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kSplit; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  SVS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, t.tmap_b[0], t.tmap_b[1], t.tmap_b[2], t.tmap_b[3], p));
  return SVS_OK;
}

int ck_launch_layer(const svs_unet_plan* plan, int li, const CUtensorMap& ta, const TcParams& p, cudaStream_t st) {
  const TcLayer& t = plan->tc[li];
  const bool tf32 = plan->precision == SVS_PRECISION_TF32;
#define SVS_CK_CASE(S)                                                                     \
  if (p.split_k == S)                                                                      \
    return tf32 ? launch_ck<float, true, S>(ta, t, p, st) : launch_ck<__nv_bfloat16, false, S>(ta, t, p, st);
  SVS_CK_CASE(2)
  SVS_CK_CASE(4)
  SVS_CK_CASE(8)
#undef SVS_CK_CASE
  return fail(SVS_ERR_NOT_IMPLEMENTED, "ck_launch_layer: unsupported split");
}

// =================================================================================================
// (2) TMA multicast
// =================================================================================================
// kClusterN = 2: 2 x 2 cluster (A and B shared); 1: 2 x 1 cluster (B shared)
template <typename OutT, bool kTf32, int kClusterN>
__global__ void __launch_bounds__(kTcThreads)
tc_conv_mc_kernel(const __grid_constant__ CUtensorMap tmap_a_half, const __grid_constant__ CUtensorMap tmap_b0,
                  const __grid_constant__ CUtensorMap tmap_b1, const __grid_constant__ CUtensorMap tmap_b2,
                  const __grid_constant__ CUtensorMap tmap_b3, const TcParams p) {
  constexpr int kBlockN = kMcBlockN, kStages = kMcStages, kSwz = 128;
  constexpr int kABytes = 128 * kSwz, kBBytes = kBlockN * kSwz, kStageBytes = kABytes + kBBytes;
  constexpr int kAHalf = kABytes / 2, kBHalf = kBBytes / 2;
  constexpr int kAccCols = kBlockN, kTmemCols = 2 * kAccCols;
  constexpr int kClusterSize = 2 * kClusterN;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * kStageBytes + 8 * (2 * kStages + 4));
  float* sbias = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes + 256);
  for (int i = threadIdx.x; i < p.cout; i += kTcThreads) sbias[i] = __ldg(&p.bias[i]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int mi = rank & 1, ni = (kClusterN == 2) ? (rank >> 1) : 0;
  // who receives my A half (my M tile's CTAs) / my B half (my N tile's CTAs) / who writes into my stages
  const uint16_t mask_a = (kClusterN == 2) ? static_cast<uint16_t>((1u << rank) | (1u << (rank ^ 2))) : 0;
  const uint16_t mask_b = static_cast<uint16_t>((1u << rank) | (1u << (rank ^ 1)));
  const uint16_t mask_release = static_cast<uint16_t>(mask_a | mask_b);

  // tiles: cluster slot q = (z, n-tile group, m-tile pair); the cluster's CTAs take (2 mp + mi, kClusterN ng + ni)
  const int m_pairs = p.m_tiles / 2, n_groups = p.n_tiles / kClusterN;
  const int per_z = m_pairs * n_groups;
  const int total_slots = per_z * p.n_phases * p.split_k;
  const int n_clusters = gridDim.x / kClusterSize;
  const int cluster_id = blockIdx.x / kClusterSize;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kClusterN == 2 ? 3 : 2);       // releases from every CTA that writes into this stage
    }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 4); }
    fence_barrier_init();
    tma_prefetch_desc(&tmap_a_half);
    tma_prefetch_desc(&tmap_b0);
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // partners' barriers are initialised before anything is multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();

  auto decode = [&](int slot, int& z, int& mt, int& nt) {
    z = slot / per_z;
    const int rem = slot - z * per_z;
    const int ng = rem / m_pairs, mp = rem - ng * m_pairs;
    mt = 2 * mp + mi;
    nt = kClusterN * ng + ni;
  };
  auto k_range = [&](int z, int& phase, int& c_begin, int& n_iter) {
    phase = z / p.split_k;
    const int split = z - phase * p.split_k;
    const int n_chunks = p.n_chunks[phase];
    const int per_split = (n_chunks + p.split_k - 1) / p.split_k;
    c_begin = split * per_split;
    n_iter = max(0, min(n_chunks, c_begin + per_split) - c_begin);
  };

  if (warp == 0) {
    // ===== TMA producer =====
    int it = 0;
    for (int slot = cluster_id; slot < total_slots; slot += n_clusters) {
      int z, mt, nt, phase, c_begin, n_iter;
      decode(slot, z, mt, nt);
      k_range(z, phase, c_begin, n_iter);
      const int tw = mt % p.ntw, th = (mt / p.ntw) % p.nth, tb = mt / (p.ntw * p.nth);
      const CUtensorMap* tb_map = phase == 0 ? &tmap_b0 : phase == 1 ? &tmap_b1 : phase == 2 ? &tmap_b2 : &tmap_b3;
      const TcChunk* chunks = p.chunks + p.chunk_begin[phase];
      const int nb_half = p.nb / 2;
      for (int i = 0; i < n_iter; ++i, ++it) {
        const int s = it % kStages;
        mbar_wait(empty_bar(s), ((it / kStages) & 1) ^ 1);
        const int ci = c_begin + i;
        const TcChunk ch = chunks[ci];
        const uint32_t a_dst = smem_base + s * kStageBytes;
        if (elect_one_sync()) {
          mbar_expect_tx(full_bar(s), kStageBytes);
          if constexpr (kClusterN == 2) {
            tma_load_5d_mc(a_dst + ni * kAHalf, &tmap_a_half, full_bar(s), ch.c_inner, tw * p.bw + ch.dw, ch.ph,
                           th * p.bh + ch.dh, tb * p.nb + ni * nb_half, mask_a);
          } else {
            tma_load_5d(a_dst, &tmap_a_half, full_bar(s), ch.c_inner, tw * p.bw + ch.dw, ch.ph, th * p.bh + ch.dh,
                        tb * p.nb);
            tma_load_5d(a_dst + kAHalf, &tmap_a_half, full_bar(s), ch.c_inner, tw * p.bw + ch.dw, ch.ph,
                        th * p.bh + ch.dh, tb * p.nb + nb_half);
          }
          tma_load_2d_mc(a_dst + kABytes + mi * kBHalf, tb_map, full_bar(s), ci * p.block_k, nt * kBlockN + mi * 64,
                         mask_b);
        }
        __syncwarp();
      }
    }
    // tail: every release aimed at this CTA's empty barriers has landed before the CTA may leave
    for (int u = max(0, it - kStages); u < it; ++u) mbar_wait(empty_bar(u % kStages), (u / kStages) & 1);
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc<kTf32, kBlockN>();
    int it = 0, t = 0;
    for (int slot = cluster_id; slot < total_slots; slot += n_clusters, ++t) {
      int z, mt, nt, phase, c_begin, n_iter;
      decode(slot, z, mt, nt);
      k_range(z, phase, c_begin, n_iter);
      const int as = t & 1;
      mbar_wait(tmem_empty_bar(as), ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * kAccCols;
      for (int i = 0; i < n_iter; ++i, ++it) {
        const int s = it % kStages;
        mbar_wait(full_bar(s), (it / kStages) & 1);
        tc_fence_after();
        const uint32_t acc0 = i > 0 ? 1u : 0u;
        dispatch_stage_mc<0, kStages>(s, [&](auto sc) {
          constexpr int S = decltype(sc)::value;
          const uint32_t a_addr = smem_base + S * kStageBytes;
          const uint64_t da = make_smem_desc<kSwz>(a_addr);
          const uint64_t db = make_smem_desc<kSwz>(a_addr + kABytes);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma<kTf32>(tmem_d, da + 2u * k, db + 2u * k, idesc, k > 0 ? 1u : acc0);
            // release this stage in every CTA that writes into it (multicast arrive on their empty barriers)
            umma_commit_mc(bar_base + 8u * (kStages + S), mask_release);
          }
          __syncwarp();
        });
      }
      if (elect_one_sync()) umma_commit(tmem_full_bar(as));
      __syncwarp();
    }
  } else {
    // ===== epilogue (identical to tc_conv_kernel) =====
    const int q = warp & 3;
    const int r = 32 * q + lane;
    const int iw = r % p.bw;
    const int ih = (r / p.bw) % p.bh;
    const int ib = r / (p.bw * p.bh);
    int t = 0;
    for (int slot = cluster_id; slot < total_slots; slot += n_clusters, ++t) {
      int z, mt, nt, phase, c_begin, n_iter;
      decode(slot, z, mt, nt);
      k_range(z, phase, c_begin, n_iter);
      const int tw = mt % p.ntw, th = (mt / p.ntw) % p.nth, tb = mt / (p.ntw * p.nth);
      const int n0 = nt * kBlockN;
      const int gx = tw * p.bw + iw, gy = th * p.bh + ih, b = tb * p.nb + ib;
      const bool valid = b < p.batch;
      const int as = t & 1;
      mbar_wait(tmem_full_bar(as), (t >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * kAccCols;
      if (p.split_k == 1) {
        OutT* const out_base = reinterpret_cast<OutT*>(p.out);
        const int oy = gy * p.out_scale + p.py[phase], ox = gx * p.out_scale + p.px[phase];
        OutT* dst0 = out_base + ((static_cast<size_t>(b) * p.hout + oy) * p.wout + ox) * p.out_pitch + p.out_coff + n0;
#pragma unroll 2
        for (int c = 0; c < kBlockN; c += 32) {
          uint32_t v[32];
          if (n_iter > 0) {
            tmem_ld16(taddr + c, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
            tmem_ld16(taddr + c + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
          }
#pragma unroll
          for (int h = 0; h < 32; h += 16) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(&sbias[n0 + c + h + i]);
              f[i] = tc_act(__uint_as_float(v[h + i]) + bv.x, p.act);
              f[i + 1] = tc_act(__uint_as_float(v[h + i + 1]) + bv.y, p.act);
              f[i + 2] = tc_act(__uint_as_float(v[h + i + 2]) + bv.z, p.act);
              f[i + 3] = tc_act(__uint_as_float(v[h + i + 3]) + bv.w, p.act);
            }
            if constexpr (kTf32) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = round_tf32(f[i]);
            }
            if (valid) store16(dst0 + c + h, f);
          }
        }
      } else {
        float* dst = p.partial +
                     (static_cast<size_t>(z) * p.m_pad + static_cast<size_t>(mt) * 128 + r) * p.cout + n0;
#pragma unroll 1
        for (int c = 0; c < kBlockN; c += 16) {
          uint32_t v[16];
          float f[16];
          if (n_iter > 0) {
            tmem_ld16(taddr + c, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0u;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          store16(dst + c, f);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(as));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // no CTA leaves while a partner may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
bool mc_supported(const svs_unet_plan* plan, int li, int batch) {
  static const bool off = [] { const char* e = std::getenv("SVS_NO_MULTICAST"); return e && e[0] == '1'; }();
  if (off) return false;
  const TcLayer& t = plan->tc[li];
  const LayerGeom& g = kLayers[li];
  if (!t.enabled || t.merged || t.swz != 128 || t.block_n != 128 || t.nb < 2 || (t.nb & 1)) return false;
  int m_tiles, split;
  tc_tiling(t, g, batch, &m_tiles, &split);
  if (m_tiles & 1) return false;
  (void)split;
  return true;
}

template <typename OutT, bool kTf32, int kClusterN>
static int launch_mc(const CUtensorMap& ta, const TcLayer& t, const TcParams& p, int total_slots, cudaStream_t st) {
  auto kern = tc_conv_mc_kernel<OutT, kTf32, kClusterN>;
  SVS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMcSmemBytes)));
  constexpr int kCluster = 2 * kClusterN;
  int clusters = num_sms() / kCluster;
  if (clusters > total_slots) clusters = total_slots;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * kCluster);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = kMcSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  SVS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, t.tmap_b_half[0], t.tmap_b_half[1], t.tmap_b_half[2],
                                  t.tmap_b_half[3], p));
  return SVS_OK;
}

// `p` is the fully populated TcParams of tc_launch_layer
int mc_launch_layer(const svs_unet_plan* plan, int li, const Workspace& ws, int batch, const TcParams& p,
                    cudaStream_t st) {
  const TcLayer& t = plan->tc[li];
  const LayerGeom& g = kLayers[li];
  const bool tf32 = plan->precision == SVS_PRECISION_TF32;
  const int es = plan->elem_size;
  CUtensorMap ta;
  {
    // the A map of conv_tc.cu with HALF the images per box (the box's outermost extent)
    const cuuint64_t ct = kBufGeom[g.in_buf].c, H = g.hin, W = g.win;
    cuuint64_t dims[5], strides[4];
    if (!g.transposed) {
      dims[0] = 2 * ct; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = batch;
      strides[0] = 2 * ct * es; strides[1] = W * ct * es; strides[2] = 2 * W * ct * es; strides[3] = H * W * ct * es;
    } else {
      dims[0] = ct; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = batch;
      strides[0] = ct * es; strides[1] = W * ct * es; strides[2] = W * ct * es; strides[3] = H * W * ct * es;
    }
    const cuuint32_t box[5] = {static_cast<cuuint32_t>(t.block_k), static_cast<cuuint32_t>(t.bw), 1,
                               static_cast<cuuint32_t>(t.bh), static_cast<cuuint32_t>(t.nb / 2)};
    int rc = encode_tensor_map(&ta, tf32, 5, ws.buf[g.in_buf], dims, strides, box, 128);
    if (rc != SVS_OK) return rc;
  }
  const int n_tiles = p.n_tiles;
  const int cluster_n = (n_tiles % 2 == 0) ? 2 : 1;
  const int total_slots = (p.m_tiles / 2) * (n_tiles / cluster_n) * t.n_phases * p.split_k;
  if (cluster_n == 2)
    return tf32 ? launch_mc<float, true, 2>(ta, t, p, total_slots, st)
                : launch_mc<__nv_bfloat16, false, 2>(ta, t, p, total_slots, st);
  return tf32 ? launch_mc<float, true, 1>(ta, t, p, total_slots, st)
              : launch_mc<__nv_bfloat16, false, 1>(ta, t, p, total_slots, st);
}

}  // namespace svs
