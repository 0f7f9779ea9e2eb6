// K3/K4-zc: 5x5 stride-2 convolution and transposed convolution as a ZERO-COPY implicit GEMM on tcgen05.
//
// tc_conv_kernel (conv_tc.cu) lets TMA do the im2col: every tap re-loads its own 128-pixel slab, so
// a conv moves 25/4 x and a phase-merged deconv 9 x its input through L2 -> SMEM, and every mid layer
// ends up bound by that traffic.  Here each 18 x 10 pixel HALO SLAB (128-byte rows: 64 bf16 / 32 fp32
// channels, 128B swizzle) of the 16 x 8 pixel tile is landed in shared memory ONCE by one TMA box, and
// the taps are *descriptor offsets* into it: for a tap shifted by (dy, dx) the A operand starts at row
// ((dy+1)*10 + (dx+1)) of the slab and its sixteen 8-row core-matrix groups (one image row each) are
// 10 rows = 1280 B apart — exactly the UMMA descriptor's stride-byte-offset.  The 128B-swizzle XOR is
// a function of the absolute shared-memory address bits (measured: base_offset = 0 reproduces TMA's
// placement for any 128-byte-aligned start), so shifted operands read back what TMA wrote.
//
//   conv   (stride 2): slabs are parity planes of the rank-5 view (pw*Ct + c, W/2, ph, H/2, B); tap
//          (kh, kw) = (2 dy + ph + 2, 2 dx + pw + 2).  A 128-byte row may span both column parities
//          (conv2: Ct = 32) or both halves of a concat buffer; 32-byte k-steps whose weights are all
//          zero (the other half of the concat buffer, kw = 5) are SKIPPED, not multiplied.
//   deconv (phase merged): slabs are 64-channel slices of (c, W, 1, H, B); N = 4 phases x Cout and
//          phase (py, px) uses tap (kh, kw) = (py + 2 - 2 dy, px + 2 - 2 dx) when it exists.
//   weights: one [N][128 B] K-major chunk per tap; RESIDENT in shared memory for the whole persistent
//          CTA when they fit (conv2, deconv5), otherwise streamed through a TMA ring.
#include "unet_internal.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>
#include <vector>

namespace svs {

extern long long* g_tc_dbg;
extern int g_tc_dbg_layer;
int zc_launch_layer_io(const svs_unet_plan* plan, int li, const ZcIo& io, int batch, cudaStream_t st);

constexpr int kZcThreads = 192;
constexpr int kZcBw = 8, kZcBh = 16;                 // M tile: 16 image rows x 8 pixels
constexpr int kZcPw = kZcBw + 2, kZcPh = kZcBh + 2;  // halo slab 18 x 10 rows
// kRow = bytes of one slab row (= TMA / UMMA swizzle span): 128 by default; 64 / 32 when the layer reads a channel
// window narrower than 128 bytes (conv3's 32 skip channels of a 64-channel concat pixel, conv2's 16 of 32) so that
// the other half of the concat buffer is neither fetched from DRAM nor carried through L2 -> shared memory.
template <int kRow> __host__ __device__ constexpr int zc_a_tx() { return kZcPw * kZcPh * kRow; }            // bytes landed per slab
template <int kRow> __host__ __device__ constexpr int zc_a_slot() { return (zc_a_tx<kRow>() + 1023) / 1024 * 1024; }   // slot pitch

struct ZcParams {
  ZcSchedule sch;
  int row_elems;                 // elements per 128-byte row
  int ntw, nth, m_tiles, batch;
  void* out;
  int out_pitch, out_coff, hout, wout, out_scale;
  const float* bias;
  int cout_phase, merged, act;
  int resident;                  // weights resident in smem
  int keep_fp32;                 // fp32 outputs are NOT rounded to TF32 (training: pre-BatchNorm z)
  int wait_first;                // griddepcontrol.wait before the resident-weight preload (weights repacked per step)
  long long* dbg;                // profiling: per CTA 64 clock64 stamps
};

// kStoreBytes: staging ring of the TMA-store epilogue (0 = per-thread global stores)
template <int kBlockN, int kASlots, int kBSlots, int kRow, int kStoreBytes = 0>
constexpr size_t zc_smem_bytes() {
  return static_cast<size_t>(kASlots) * zc_a_slot<kRow>() + static_cast<size_t>(kBSlots) * kBlockN * kRow + kStoreBytes +
         1024 + 1024 + 2048;
}

__device__ __forceinline__ float zc_act(float v, int act) {
  if (act == ACT_LEAKY) return v > 0.0f ? v : 0.2f * v;
  return fmaxf(v, 0.0f);
}
__device__ __forceinline__ void zc_store16(__nv_bfloat16* dst, const float (&f)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  reinterpret_cast<uint4*>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
  reinterpret_cast<uint4*>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ void zc_store16(float* dst, const float (&f)[16], int keep_fp32 = 0) {   // fp32 activations feed kind::tf32
#pragma unroll
  for (int i = 0; i < 4; ++i)
    reinterpret_cast<float4*>(dst)[i] = keep_fp32 ? make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3])
                                                  : make_float4(round_tf32(f[4 * i]), round_tf32(f[4 * i + 1]),
                                                                round_tf32(f[4 * i + 2]), round_tf32(f[4 * i + 3]));
}
__device__ __forceinline__ void zc_store16(__nv_bfloat16* dst, const float (&f)[16], int) { zc_store16(dst, f); }

// ---- compile-time tap tables ---------------------------------------------------------------------
// The MMA issuer must not do per-tap address arithmetic in the vector datapath: tcgen05.mma takes its
// descriptors from UNIFORM registers and every R2UR hop costs tens of cycles, which at 4 MMAs per tap
// dominated the issue rate (measured 178 cycles per N=64 MMA against a 48-cycle hardware floor).  With the
// tap set known at compile time the issue loop unrolls into straight-line code whose descriptors are
// "slab base + constant".  kmask(s, dy, dx) = active 32-byte k-steps of tap (dy, dx) of slab s, 0 = no tap;
// the order (s, dy, dx ascending, inactive skipped) is the order zc_plan_layer() packs the weights in.
struct ZcRuntimeTaps { static constexpr bool kStatic = false; static constexpr int kSlabs = 0; static constexpr bool kMerged = false;
  __host__ __device__ static constexpr int kmask(int, int, int) { return 0; } };
template <int kNSlabs> struct ZcDeconvTaps {      // phase-merged deconv: every slab uses all 9 shifts, full K
  static constexpr bool kStatic = true; static constexpr int kSlabs = kNSlabs; static constexpr bool kMerged = true;
  __host__ __device__ static constexpr int kmask(int, int, int) { return 0xF; } };
struct ZcConv2Taps {                              // Ct = 32: row = [pw0: d-half | skip | pw1: d-half | skip]; slab = ph
  static constexpr bool kStatic = true; static constexpr int kSlabs = 2; static constexpr bool kMerged = false;
  __host__ __device__ static constexpr int kmask(int s, int dy, int dx) { return (2 * dy + s + 2 > 4) ? 0 : (dx <= 0 ? 0xA : 0x2); } };
struct ZcConvParity2Taps {                        // two 128-byte channel windows per column parity (TF32 conv4):
  static constexpr bool kStatic = true; static constexpr int kSlabs = 8; static constexpr bool kMerged = false;   // slab = ph*4 + pw*2 + window
  __host__ __device__ static constexpr int kmask(int s, int dy, int dx) {
    return (2 * dy + (s >> 2) + 2 > 4 || 2 * dx + ((s >> 1) & 1) + 2 > 4) ? 0 : 0xF; } };
template <int kKMask> struct ZcConvParityTaps {   // slabs (ph, pw) = (s >> 1, s & 1): conv3 (skip half = 0xC), conv4 (0xF)
  static constexpr bool kStatic = true; static constexpr int kSlabs = 4; static constexpr bool kMerged = false;
  __host__ __device__ static constexpr int kmask(int s, int dy, int dx) {
    return (2 * dy + (s >> 1) + 2 > 4 || 2 * dx + (s & 1) + 2 > 4) ? 0 : kKMask; } };

// kBSlots: ring depth when streaming, number of taps when resident
__device__ __forceinline__ uint32_t zc_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void zc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load delivered to the same shared-memory offset (and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// kCluster = 2 (experiment, SVS_ZC_MCAST=1): the two CTAs of a cluster walk neighbouring tiles in lock step and SHARE
// the streamed weight chunks: each CTA issues one half of every chunk as a multicast TMA load that lands in both
// CTAs.  A slot may be refilled only when BOTH CTAs have consumed it: the MMA warp's commit arrives on the slot's
// `empty` barrier of both CTAs (count 2).  Parity-green but not faster (see zc_launch_layer), so off by default.
// kTrim (phase-merged deconvs with streamed weights): a tap (dy, dx) only reaches the sub-pixel phases whose kernel
// index kh = py + 2 - 2 dy, kw = px + 2 - 2 dx exists, i.e. py = 0 alone when dy = -1 and px = 0 alone when dx = -1.
// With the accumulator columns ordered [(0,0) (0,1) (1,0) (1,1)] x Cout, the nine taps need 1, 2, 2 (as two
// single-phase MMAs) or 4 phase blocks instead of 4 each: 25 blocks of weights instead of 36 are fetched (the layers
// are bound by bytes into shared memory) and 25 / 36 of the MMA columns are issued.  The tile's very first tap still
// runs over all four blocks (its unused ones hold zero weights) so that one instruction initialises every column.
// kDual: the CTA works on TWO neighbouring M tiles at once (two accumulators per TMEM stage, a slab slot holds both
// tiles' halo slabs) and every streamed weight chunk feeds both: the layer is bound by bytes streamed into the SM, and
// this halves the weight bytes per MMA without a cluster.  N <= 128 (4 x N TMEM columns).
// kStoreUnits > 0: TMA-STORE EPILOGUE.  The per-thread stores of the plain epilogue are 16-byte pieces at one pixel
// pitch per lane: 32 cache lines = 32 LSU wavefronts per instruction, 2,048 per 128 x 128 tile, which is what the
// 3.6k-cycle exposed epilogue of the one-tile-per-CTA layers was.  Here a thread writes its 32 accumulator columns
// into a shared-memory unit laid out as the TMA box(es) of the step (rows of 32 / 64 / 128 bytes in the matching
// TMA swizzle: row-per-lane 16-byte stores are conflict free in all three), and one elected thread hands the unit to
// the TMA engine (cp.async.bulk.tensor.5d.global.shared: UTMASTG).  Units form a ring of kStoreUnits; the issuing
// thread waits for the store kStoreUnits - 1 steps back to have READ its unit before the barrier that lets the
// others fill the next one, so there is one 128-thread barrier per 32 columns.  The output is addressed through a
// 5-D map (channel, px, x, py, batch * grid rows + y): conv layers have px = py = 1, merged deconvs store one
// sub-pixel phase per box.  kStoreAlias (launches with ONE tile per CTA, i.e. batch 64): the ring lives in the slab
// slots, which are dead once the tile's MMAs have retired, so the operand rings keep their full depth.
template <typename OutT, bool kTf32, int kBlockN, int kASlots, int kBSlots, typename Taps, int kRow = 128, int kCluster = 1,
          bool kTrim = false, bool kDual = false, int kStoreUnits = 0, bool kStoreAlias = false>
__global__ void __launch_bounds__(kZcThreads)
zc_conv_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ ZcParams p) {
  static_assert(kCluster == 1 || kCluster == 2, "CTA pairs only");
  static_assert(!kTrim || (kCluster == 1 && kRow == 128 && Taps::kStatic && kBlockN % 64 == 0), "trimmed form");
  static_assert(!kDual || (kCluster == 1 && !kTrim && Taps::kStatic && kBlockN <= 128), "dual-tile form");
  constexpr int kTiles = kDual ? 2 : 1;
  constexpr uint16_t kMask = (1u << kCluster) - 1u;
  constexpr int kBBytes = kBlockN * kRow;
  constexpr int kZcATx = kTiles * zc_a_tx<kRow>(), kZcASlot = kTiles * zc_a_slot<kRow>(), kZcATile = zc_a_slot<kRow>();
  constexpr int kKSteps = kRow / 32;                              // UMMA K = 32 bytes
  constexpr uint64_t kLayout = kRow == 128 ? 2 : (kRow == 64 ? 4 : 6);   // UMMA layout type of the swizzle span
  constexpr int kAccCols = kBlockN < 32 ? 32 : kBlockN;
  constexpr int kStageCols = kTiles * kAccCols;                   // accumulator columns of one TMEM stage
  constexpr int kTmemCols = 2 * kStageCols;
  constexpr int kNBar = 2 * kBSlots + 2 * kASlots + 4;
  constexpr int kUnitBytes = 128 * 32 * static_cast<int>(sizeof(OutT));     // one 32-column step of the tile
  constexpr int kStoreBytes = kStoreAlias ? 0 : kStoreUnits * kUnitBytes;
  static_assert(kStoreUnits == 0 || (kStoreUnits >= 2 && !kDual && kBlockN % 32 == 0), "store ring");
  static_assert(!kStoreAlias || kStoreUnits * kUnitBytes <= kASlots * kZcASlot, "aliased store ring fits the slab slots");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_base = smem_base + kASlots * kZcASlot;
  const uint32_t store_base = kStoreAlias ? smem_base : b_base + kBSlots * kBBytes;   // 1024-byte aligned: slots and chunks are
  const size_t bar_off = static_cast<size_t>(kASlots) * kZcASlot + static_cast<size_t>(kBSlots) * kBBytes + kStoreBytes;
  const uint32_t bar_base = smem_base + static_cast<uint32_t>(bar_off);
  auto full_b = [&](int s) { return bar_base + 8u * s; };
  auto empty_b = [&](int s) { return bar_base + 8u * (kBSlots + s); };
  auto full_a = [&](int s) { return bar_base + 8u * (2 * kBSlots + s); };
  auto empty_a = [&](int s) { return bar_base + 8u * (2 * kBSlots + kASlots + s); };
  auto tmem_full = [&](int s) { return bar_base + 8u * (2 * kBSlots + 2 * kASlots + s); };
  auto tmem_empty = [&](int s) { return bar_base + 8u * (2 * kBSlots + 2 * kASlots + 2 + s); };
  static_assert(kNBar * 8 + 8 <= 1024, "barrier area");
  const uint32_t tmem_slot = bar_base + 8u * kNBar;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * kNBar);
  float* sbias = reinterpret_cast<float*>(smem_gen + bar_off + 1024);
  for (int i = threadIdx.x; i < p.cout_phase; i += kZcThreads) sbias[i] = __ldg(&p.bias[i]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = p.dbg ? p.dbg + 64 * blockIdx.x : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  const int total_tiles = p.m_tiles / kTiles;         // work units (kBlockN covers all of N; dual: tile pairs)
  const int n_slabs = p.sch.n_slabs, n_taps = p.sch.n_taps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBSlots; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), kCluster); }
    for (int s = 0; s < kASlots; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full(s), 1); mbar_init(tmem_empty(s), 4); }
    fence_barrier_init();
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if constexpr (kStoreUnits > 0) tma_prefetch_desc(&tmap_o);
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster > 1) zc_cluster_sync();    // the partner's barriers are initialised before anything lands in them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===== TMA producer: weights (once, or a ring) and halo slabs (a ring running ahead across tiles) =====
    {
      if (p.wait_first) pdl_wait();                 // the weights were repacked by a kernel earlier in this stream
      if (p.resident) {
        if (elect_one_sync()) {
          mbar_expect_tx(full_b(0), static_cast<uint32_t>(n_taps) * kBBytes);
          for (int t = 0; t < n_taps; ++t)
            tma_load_2d(b_base + t * kBBytes, &tmap_b, full_b(0), t * p.row_elems, 0);
        }
        __syncwarp();
      }
      pdl_wait();                                 // activations of the previous layer from here on
      // slab jobs are numbered across tiles; the slab of job j+1 is requested BEFORE the weight chunks
      // of job j so the halo loads run one slab ahead of the MMAs
      auto issue_a = [&](int ja, int unit, int s) {
        const int slot = ja % kASlots;
        mbar_wait(empty_a(slot), ((ja / kASlots) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(full_a(slot), kZcATx);
#pragma unroll
          for (int u = 0; u < kTiles; ++u) {
            const int tile = unit * kTiles + u;
            const int tw = tile % p.ntw, th = (tile / p.ntw) % p.nth, tb = tile / (p.ntw * p.nth);
            tma_load_5d(smem_base + slot * kZcASlot + u * kZcATile, &tmap_a, full_a(slot), p.sch.slab_c[s], tw * kZcBw - 1,
                        p.sch.slab_ph[s], th * kZcBh - 1, tb);
          }
        }
        __syncwarp();
        if (dbg && ja < 8 && lane == 0) dbg[8 + ja] = clock64();
      };
      int ja = 0, jb = 0;
      if (static_cast<int>(blockIdx.x) < total_tiles) issue_a(0, blockIdx.x, 0);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int next_tap = 0;
        for (int s = 0; s < n_slabs; ++s, ++ja) {
          if (s + 1 < n_slabs) issue_a(ja + 1, tile, s + 1);
          else if (tile + static_cast<int>(gridDim.x) < total_tiles) issue_a(ja + 1, tile + gridDim.x, 0);
          if (!p.resident) {
            for (; next_tap < n_taps && p.sch.tap_slab[next_tap] == s; ++next_tap, ++jb) {
              const int bs = jb % kBSlots;
              mbar_wait(empty_b(bs), ((jb / kBSlots) & 1) ^ 1);
              if (elect_one_sync()) {
                if constexpr (kTrim) {               // tmap_b's box is one phase block (kBlockN / 4 rows)
                  const bool full = next_tap == 0;
                  const int pyn = (full || p.sch.tap_dy[next_tap] >= 0) ? 2 : 1;
                  const int pxn = (full || p.sch.tap_dx[next_tap] >= 0) ? 2 : 1;
                  mbar_expect_tx(full_b(bs), static_cast<uint32_t>(pyn * pxn) * (kBBytes / 4));
                  for (int py = 0; py < pyn; ++py)
                    for (int px = 0; px < pxn; ++px)
                      tma_load_2d(b_base + bs * kBBytes + (2 * py + px) * (kBBytes / 4), &tmap_b, full_b(bs),
                                  next_tap * p.row_elems, (2 * py + px) * (kBlockN / 4));
                } else if constexpr (kCluster == 1) {
                  mbar_expect_tx(full_b(bs), kBBytes);
                  tma_load_2d(b_base + bs * kBBytes, &tmap_b, full_b(bs), next_tap * p.row_elems, 0);
                } else {
                  mbar_expect_tx(full_b(bs), kBBytes);     // my half of the rows, into both CTAs (box = kBlockN / 2 rows)
                  const int half = static_cast<int>(zc_cluster_rank());
                  tma_load_2d_mcast(b_base + bs * kBBytes + half * (kBBytes / 2), &tmap_b, full_b(bs), next_tap * p.row_elems,
                                    half * (kBlockN / 2), kMask);
                }
              }
              __syncwarp();
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    {
      constexpr uint32_t idesc = make_idesc<kTf32, kBlockN>();
      int ja = 0, jb = 0, t = 0;
      if (p.resident) { mbar_wait(full_b(0), 0); tc_fence_after(); }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
        const int as = t & 1;
        mbar_wait(tmem_empty(as), ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kStageCols;
        if constexpr (Taps::kStatic) {
          // ---- straight-line issue: slabs, shifts and k-steps unrolled at compile time ----
          bool first_mma = true;
          int tap_c = 0;                                // compile-time after unrolling
#pragma unroll
          for (int s = 0; s < Taps::kSlabs; ++s, ++ja) {
            const int slot = ja % kASlots;
            mbar_wait(full_a(slot), (ja / kASlots) & 1);
            if (dbg && ja < 8 && lane == 0) dbg[16 + ja] = clock64();
            tc_fence_after();
            const uint32_t a_base = smem_base + slot * kZcASlot;
            const uint64_t da0 = static_cast<uint64_t>((a_base & 0x3FFFF) >> 4) | (1ull << 16) |
                                 (static_cast<uint64_t>((kZcPw * kRow) >> 4) << 32) | (1ull << 46) | (kLayout << 61);
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
              for (int dx = -1; dx <= 1; ++dx) {
                const int kmask = Taps::kmask(s, dy, dx);
                if (kmask == 0) continue;
                uint32_t b_addr;
                int bs = 0;
                if (p.resident) {
                  b_addr = b_base + tap_c * kBBytes;
                } else {
                  bs = jb % kBSlots;
                  mbar_wait(full_b(bs), (jb / kBSlots) & 1);
                  tc_fence_after();
                  b_addr = b_base + bs * kBBytes;
                }
                const uint64_t db = make_smem_desc<kRow>(b_addr);
                const uint32_t a_off16 = static_cast<uint32_t>(((dy + 1) * kZcPw + (dx + 1)) * kRow) >> 4;
                if (elect_one_sync()) {
                  if constexpr (kTrim) {
                    constexpr int C = kBlockN / 4;
                    constexpr uint32_t idesc1 = make_idesc<kTf32, C>(), idesc2 = make_idesc<kTf32, 2 * C>();
                    constexpr uint32_t blk16 = static_cast<uint32_t>(C * kRow) >> 4;     // one phase block of B rows
                    const bool full = (s == 0 && dy == -1 && dx == -1) || (dy >= 0 && dx >= 0);
#pragma unroll
                    for (int k = 0; k < kKSteps; ++k) {
                      const uint64_t dak = da0 + a_off16 + 2u * k, dbk = db + 2u * k;
                      if (full) {
                        umma<kTf32>(tmem_d, dak, dbk, idesc, first_mma ? 0u : 1u);
                      } else if (dy == -1 && dx == -1) {                // phase (0,0)
                        umma<kTf32>(tmem_d, dak, dbk, idesc1, 1u);
                      } else if (dy == -1) {                            // phases (0,0) (0,1)
                        umma<kTf32>(tmem_d, dak, dbk, idesc2, 1u);
                      } else {                                          // dx == -1: phases (0,0) and (1,0)
                        umma<kTf32>(tmem_d, dak, dbk, idesc1, 1u);
                        umma<kTf32>(tmem_d + 2 * C, dak, dbk + 2u * blk16, idesc1, 1u);
                      }
                      first_mma = false;
                    }
                  } else {
#pragma unroll
                  for (int k = 0; k < kKSteps; ++k) {
                    if (kmask & (1 << k)) {
                      umma<kTf32>(tmem_d, da0 + a_off16 + 2u * k, db + 2u * k, idesc, first_mma ? 0u : 1u);
                      if constexpr (kDual)     // the second tile's slab sits kZcATile bytes further in the same slot
                        umma<kTf32>(tmem_d + kAccCols, da0 + (static_cast<uint32_t>(kZcATile) >> 4) + a_off16 + 2u * k,
                                    db + 2u * k, idesc, first_mma ? 0u : 1u);
                      first_mma = false;
                    }
                  }
                  }
                  if (!p.resident) {
                    if constexpr (kCluster == 1) umma_commit(empty_b(bs));
                    else umma_commit_mcast(empty_b(bs), kMask);
                  }
                }
                __syncwarp();
                first_mma = false;
                if (!p.resident) ++jb;
                ++tap_c;
              }
            }
            if (elect_one_sync()) umma_commit(empty_a(slot));
            __syncwarp();
          }
        } else {
        uint32_t first = 0;                             // accumulate flag: 0 for the tile's first MMA
        int tap = 0;
        for (int s = 0; s < n_slabs; ++s, ++ja) {
          const int slot = ja % kASlots;
          mbar_wait(full_a(slot), (ja / kASlots) & 1);
          if (dbg && ja < 8 && lane == 0) dbg[16 + ja] = clock64();
          tc_fence_after();
          const uint32_t a_base = smem_base + slot * kZcASlot;
          for (; tap < n_taps && p.sch.tap_slab[tap] == s; ++tap) {
            uint32_t b_addr;
            int bs = 0;
            if (p.resident) {
              b_addr = b_base + tap * kBBytes;
            } else {
              bs = jb % kBSlots;
              mbar_wait(full_b(bs), (jb / kBSlots) & 1);
              tc_fence_after();
              b_addr = b_base + bs * kBBytes;
            }
            const uint32_t a_addr = a_base + ((p.sch.tap_dy[tap] + 1) * kZcPw + (p.sch.tap_dx[tap] + 1)) * kRow;
            const uint32_t sbo = kZcPw * kRow;
            // K-major; 8-row groups (one image row) are 10 slab rows apart; base_offset 0
            const uint64_t da = static_cast<uint64_t>((a_addr & 0x3FFFF) >> 4) | (1ull << 16) |
                                (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46) | (kLayout << 61);
            const uint64_t db = make_smem_desc<kRow>(b_addr);
            const int kmask = p.sch.tap_kmask[tap];
            const int k_lo = __ffs(kmask) - 1;             // first active k-step of this tap
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < kKSteps; ++k) {
                if (kmask & (1 << k)) umma<kTf32>(tmem_d, da + 2u * k, db + 2u * k, idesc, (first != 0u || k != k_lo) ? 1u : 0u);
              }
              if (!p.resident) {
                if constexpr (kCluster == 1) umma_commit(empty_b(bs));
                else umma_commit_mcast(empty_b(bs), kMask);
              }
            }
            __syncwarp();
            first = 1u;
            if (!p.resident) ++jb;
          }
          if (elect_one_sync()) umma_commit(empty_a(slot));
          __syncwarp();
        }
        }
        if (elect_one_sync()) umma_commit(tmem_full(as));
        __syncwarp();
        if (dbg && t < 8 && lane == 0) dbg[24 + t] = clock64();
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue =====
    if constexpr (kStoreUnits == 0) pdl_wait();   // before the first global store
    const int q = warp & 3;
    const int r = 32 * q + lane;
    const int ix = r & 7, iy = r >> 3;
    int t = 0;
    int store_step = 0;                           // TMA-store epilogue: 32-column steps since the kernel started
    OutT* const out_base = reinterpret_cast<OutT*>(p.out);
    // max(v, slope * v): LeakyReLU (0.2), ReLU (0) or the identity (1: the training forward stores the raw sums)
    const float slope = p.act == ACT_LEAKY ? 0.2f : (p.act == ACT_RELU ? 0.0f : 1.0f);
    const int cp_log2 = 31 - __clz(p.cout_phase), cp_mask = p.cout_phase - 1;   // channels per phase: a power of two
    const uint32_t out_pitch = p.out_pitch, out_coff = p.out_coff;
    if constexpr (kStoreUnits > 0) {
      // ---- TMA-store epilogue (see the kernel comment) ----
      constexpr int kEs = static_cast<int>(sizeof(OutT));
      constexpr int kPhaseCols = Taps::kMerged ? kBlockN / 4 : kBlockN;      // contiguous channels of one pixel
      constexpr int kBoxCols = kPhaseCols < 32 ? kPhaseCols : 32;            // channels per TMA box
      constexpr int kBoxW = kBoxCols * kEs;                                  // bytes per box row: 32 / 64 / 128
      constexpr int kBoxes = 32 / kBoxCols;                                  // boxes per 32-column step
      static_assert(kBoxW == 32 || kBoxW == 64 || kBoxW == 128, "box row = one TMA swizzle span");
      static_assert(kStoreUnits <= 4, "one issuing warp per unit");
      constexpr int kChunks = kBoxW / 16;
      // XOR of the 16-byte chunk index inside a box row: TMA's swizzle is a function of the shared-memory address
      const uint32_t swz = kBoxW == 128 ? (r & 7) : kBoxW == 64 ? ((r >> 1) & 3) : ((r >> 2) & 1);
      // unit u is stored (and its bulk groups are waited for) by lane 0 of epilogue warp u: the ~150 cycles of
      // coordinate set-up + issue rotate over the four warps instead of holding up one of them at every barrier
      const int my_unit = lane == 0 ? q : -1;
      // Pass -1 is a DRY RUN (no tile, nothing stored, nothing signalled) made while the MMAs of the first tile are
      // still running: the epilogue is ~2 KB of straight-line code that would otherwise be fetched cold at the one
      // moment it is on the critical path (measured: 560 cycles before the first column and 720 instead of 350 for
      // the first 32-column step).  One loop body serves both passes so that the compiler cannot specialise it.
#pragma unroll 1
      for (int it = -1, unit = blockIdx.x; it < 0 || unit < total_tiles; ++it) {
        const bool live = it >= 0;
        const int as = it & 1;
        if (live) {
          if (it == 0) pdl_wait();                         // before the first global store
          mbar_wait(tmem_full(as), (it >> 1) & 1);
          if (dbg && it < 8 && threadIdx.x == 64) dbg[32 + it] = clock64();
          tc_fence_after();
        }
        const int tw = unit % p.ntw, th = (unit / p.ntw) % p.nth, b = unit / (p.ntw * p.nth);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * kStageCols;
#pragma unroll 1
        for (int c = 0; c < kBlockN; c += 32) {
          uint32_t v[32];
          tmem_ld16(taddr + c, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          tmem_ld16(taddr + c + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          tmem_ld_wait();
          const int cur = store_step % kStoreUnits;
          const uint32_t ubase = store_base + cur * kUnitBytes;
#pragma unroll
          for (int bx = 0; bx < kBoxes; ++bx) {
            const uint32_t row = ubase + bx * (128 * kBoxW) + r * kBoxW;
            const int ch0 = (c + bx * kBoxCols) & (kPhaseCols - 1);           // channel of the box's first column
#pragma unroll
            for (int k = 0; k < kChunks; ++k) {
              constexpr int kPer = 16 / kEs;                                   // values per 16-byte chunk
              float f[kPer];
#pragma unroll
              for (int i = 0; i < kPer; i += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(&sbias[ch0 + k * kPer + i]);
                const int vi = bx * kBoxCols + k * kPer + i;
                const float t0 = __uint_as_float(v[vi]) + bv.x, t1 = __uint_as_float(v[vi + 1]) + bv.y;
                const float t2 = __uint_as_float(v[vi + 2]) + bv.z, t3 = __uint_as_float(v[vi + 3]) + bv.w;
                f[i] = fmaxf(t0, fmaf(slope, t0, 0.0f));
                f[i + 1] = fmaxf(t1, fmaf(slope, t1, 0.0f));
                f[i + 2] = fmaxf(t2, fmaf(slope, t2, 0.0f));
                f[i + 3] = fmaxf(t3, fmaf(slope, t3, 0.0f));
              }
              uint32_t w[4];
              if constexpr (kEs == 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                  w[i] = *reinterpret_cast<uint32_t*>(&h);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = __float_as_uint(round_tf32(f[i]));
              }
              if (!kStoreAlias || live)                    // aliased ring: the dry run must not touch the slab slots
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((static_cast<uint32_t>(k) ^ swz) << 4)),
                             "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                             : "memory");
            }
          }
          if (live && c + 32 >= kBlockN) {                 // last TMEM read of the tile: the MMAs may refill the stage
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(as));
          }
          fence_proxy_async();                             // the unit is read by the async proxy
          if (my_unit == (cur + 1) % kStoreUnits) bulk_wait_group_read<0>();   // the NEXT step's unit has been read out
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (live && my_unit == cur) {
#pragma unroll
            for (int bx = 0; bx < kBoxes; ++bx) {
              const int col = c + bx * kBoxCols;
              const int ph = Taps::kMerged ? col / kPhaseCols : 0;
              tma_store_5d(&tmap_o, ubase + bx * (128 * kBoxW), static_cast<int>(out_coff) + (col & (kPhaseCols - 1)), ph & 1,
                           tw * kZcBw, ph >> 1, b * p.nth * kZcBh + th * kZcBh);
            }
            bulk_commit_group();
          }
          if (live) ++store_step;
        }
        if (live) {
          if (dbg && it < 8 && threadIdx.x == 64) dbg[40 + it] = clock64();
          unit += gridDim.x;
        }
      }
      if (my_unit >= 0) bulk_wait_group_all();             // every store of this CTA has been written before it exits
    } else
    for (int unit = blockIdx.x; unit < total_tiles; unit += gridDim.x, ++t) {
      const int as = t & 1;
      mbar_wait(tmem_full(as), (t >> 1) & 1);
      if (dbg && t < 8 && threadIdx.x == 64) dbg[32 + t] = clock64();
      tc_fence_after();
#pragma unroll 1
      for (int u = 0; u < kTiles; ++u) {
      const int tile = unit * kTiles + u;
      const int tw = tile % p.ntw, th = (tile / p.ntw) % p.nth, b = tile / (p.ntw * p.nth);
      const int gx = tw * kZcBw + ix, gy = th * kZcBh + iy;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * kStageCols + u * kAccCols;
      constexpr int kStep = kBlockN >= 32 ? 32 : 16;
      // The epilogue of tile t runs under the MMAs of tile t+1 with ONE warp per scheduler, so it is bound by
      // dependent-issue latency: keep it short (branch-free activation, shifts, 32-bit offsets) or it, not the
      // tensor pipe, sets the tile time of the small-N layers (deconv5: 36 MMAs = 1,728 cycles per tile).
      const uint32_t pix0 = static_cast<uint32_t>((b * p.hout + gy * p.out_scale) * p.wout + gx * p.out_scale);
#pragma unroll 2
      for (int c = 0; c < kBlockN; c += kStep) {
        uint32_t v[kStep];
        tmem_ld16(taddr + c, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        if constexpr (kStep == 32) tmem_ld16(taddr + c + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < kStep; h += 16) {
          int ch = c + h;
          uint32_t pix = pix0;
          if (p.merged) {                                   // 16-column groups never straddle a phase
            const int ph = ch >> cp_log2;
            ch &= cp_mask;
            pix += static_cast<uint32_t>((ph >> 1) * p.wout + (ph & 1));
          }
          OutT* dst = out_base + (static_cast<size_t>(pix) * out_pitch + out_coff + ch);
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(&sbias[ch + i]);
            const float t0 = __uint_as_float(v[h + i]) + bv.x, t1 = __uint_as_float(v[h + i + 1]) + bv.y;
            const float t2 = __uint_as_float(v[h + i + 2]) + bv.z, t3 = __uint_as_float(v[h + i + 3]) + bv.w;
            // max(v, slope * v + 0): LeakyReLU (slope 0.2) or ReLU (slope 0) without a branch
            f[i] = fmaxf(t0, fmaf(slope, t0, 0.0f));
            f[i + 1] = fmaxf(t1, fmaf(slope, t1, 0.0f));
            f[i + 2] = fmaxf(t2, fmaf(slope, t2, 0.0f));
            f[i + 3] = fmaxf(t3, fmaf(slope, t3, 0.0f));
          }
          zc_store16(dst, f, p.keep_fp32);
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty(as));
      if (dbg && t < 8 && threadIdx.x == 64) dbg[40 + t] = clock64();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster > 1) zc_cluster_sync();    // the partner's last commits / loads target this CTA's shared memory
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// weights: B[n][tap * row_elems + e] for the slab rows described by the schedule
struct ZcPackArgs {
  ZcSchedule sch;
  int row_elems, ct, in_coff, cin, cout, transposed, n_total;
};

template <typename E>
__global__ void zc_pack_weights_kernel(const float* __restrict__ w_fold /*[25][cin][cout]*/, ZcPackArgs a,
                                       E* __restrict__ out) {
  const int k_total = a.sch.n_taps * a.row_elems;
  const size_t total = static_cast<size_t>(a.n_total) * k_total;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / k_total);
    const int k = static_cast<int>(i % k_total);
    const int tap = k / a.row_elems, e = k % a.row_elems;
    const int s = a.sch.tap_slab[tap], dy = a.sch.tap_dy[tap], dx = a.sch.tap_dx[tap];
    const int abs_c = a.sch.slab_c[s] + e;
    int kh, kw, ci, co;
    if (!a.transposed) {
      const int pw = abs_c / a.ct;
      ci = abs_c % a.ct - a.in_coff;
      kh = 2 * dy + a.sch.slab_ph[s] + 2;
      kw = 2 * dx + pw + 2;
      co = n;
    } else {
      const int ph = n / a.cout;
      co = n % a.cout;
      ci = abs_c - a.in_coff;
      kh = (ph >> 1) + 2 - 2 * dy;
      kw = (ph & 1) + 2 - 2 * dx;
    }
    float v = 0.0f;
    if (kh >= 0 && kh <= 4 && kw >= 0 && kw <= 4 && ci >= 0 && ci < a.cin)
      v = w_fold[(static_cast<size_t>(kh * 5 + kw) * a.cin + ci) * a.cout + co];
    if constexpr (sizeof(E) == 2) out[i] = __float2bfloat16_rn(v);
    else out[i] = round_tf32(v);
  }
}

// host: is k-step `k` of tap (slab s, dy, dx) non-zero for any output column?
static bool kstep_active(const ZcSchedule& sch, const LayerGeom& g, int ct, int row_elems, int ksteps, int s, int dy, int dx, int k) {
  const int kel = row_elems / ksteps;
  for (int e = k * kel; e < (k + 1) * kel; ++e) {
    const int abs_c = sch.slab_c[s] + e;
    if (!g.transposed) {
      const int pw = abs_c / ct, ci = abs_c % ct - g.in_coff;
      const int kh = 2 * dy + sch.slab_ph[s] + 2, kw = 2 * dx + pw + 2;
      if (kh >= 0 && kh <= 4 && kw >= 0 && kw <= 4 && ci >= 0 && ci < g.cin) return true;
    } else {
      const int ci = abs_c - g.in_coff;
      if (ci < 0 || ci >= g.cin) continue;
      for (int ph = 0; ph < 4; ++ph) {
        const int kh = (ph >> 1) + 2 - 2 * dy, kw = (ph & 1) + 2 - 2 * dx;
        if (kh >= 0 && kh <= 4 && kw >= 0 && kw <= 4) return true;
      }
    }
  }
  return false;
}

int zc_plan_layer(svs_unet_plan* plan, int li, cudaStream_t st) {
  static const int mode = [] { const char* e = std::getenv("SVS_ZC_DISABLE"); return e ? std::atoi(e) : 0; }();
  const LayerGeom& g = kLayers[li];
  ZcLayer& z = plan->zc[li];
  z.enabled = false;
  if (mode == 1 || ((mode >> 4) >> li) & 1) return SVS_OK;
  const bool tf32 = plan->precision == SVS_PRECISION_TF32;
  const int es = plan->elem_size;
  const int gh = g.transposed ? g.hin : g.hout, gw = g.transposed ? g.win : g.wout;
  if (gh % kZcBh != 0 || gw % kZcBw != 0) return SVS_OK;
  const int n_total = g.transposed ? 4 * g.cout : g.cout;
  if (n_total != 32 && n_total != 64 && n_total != 128 && n_total != 256) return SVS_OK;
  const int ct = kBufGeom[g.in_buf].c;
  // slab row = 128 bytes of channels, or exactly the layer's own channel window when that is narrower (conv2, conv3:
  // the skip half of a concat pixel): the other half is then never fetched
  static const bool narrow_on = [] { const char* e = std::getenv("SVS_ZC_NARROW"); return !(e && e[0] == '0'); }();
  static const bool narrow_conv2 = [] { const char* e = std::getenv("SVS_ZC_NARROW"); return e && e[0] == '2'; }();
  int row_bytes = 128;
  if (narrow_on && !g.transposed && g.cin * es < 128 && (g.cin * es == 32 || g.cin * es == 64) &&
      (g.in_coff * es) % (g.cin * es) == 0 && (li == 2 || (li == 1 && narrow_conv2)))
    row_bytes = g.cin * es;          // conv3 only: conv2 is DRAM bound either way (32-byte sectors of 64-byte DRAM
                                     // atoms save nothing) and TMA issues ~5 cycles per slab row whatever its width,
                                     // so 32-byte rows were slower there (measured: 4,600 vs 3,750 cycles per tile)
  const int row = row_bytes / es;
  const int ksteps = row_bytes / 32;
  ZcSchedule sch{};
  auto add_slab = [&](int c, int ph) { sch.slab_c[sch.n_slabs] = c; sch.slab_ph[sch.n_slabs] = ph; return sch.n_slabs++; };
  if (!g.transposed) {
    // distinct 128-byte row windows of the parity-merged dim (2*Ct elements) that touch the input channels
    std::vector<int> windows;
    for (int pw = 0; pw < 2; ++pw)
      for (int c = 0; c < g.cin; ++c) {
        const int w0 = (pw * ct + g.in_coff + c) / row * row;
        bool seen = false;
        for (int v : windows) seen |= (v == w0);
        if (!seen) windows.push_back(w0);
      }
    if (static_cast<int>(windows.size()) * 2 > kZcMaxSlabs) return SVS_OK;
    for (int ph = 0; ph < 2; ++ph)
      for (int w0 : windows) add_slab(w0, ph);
  } else {
    if (g.cin % row != 0 || g.in_coff % row != 0 || g.cin / row > kZcMaxSlabs) return SVS_OK;
    for (int j = 0; j < g.cin / row; ++j) add_slab(g.in_coff + j * row, 0);
  }
  for (int s = 0; s < sch.n_slabs; ++s)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        int kmask = 0;
        for (int k = 0; k < ksteps; ++k) kmask |= kstep_active(sch, g, ct, row, ksteps, s, dy, dx, k) ? (1 << k) : 0;
        if (!kmask) continue;
        if (sch.n_taps >= kZcMaxTaps) return SVS_OK;
        sch.tap_slab[sch.n_taps] = static_cast<signed char>(s);
        sch.tap_dy[sch.n_taps] = static_cast<signed char>(dy);
        sch.tap_dx[sch.n_taps] = static_cast<signed char>(dx);
        sch.tap_kmask[sch.n_taps] = static_cast<signed char>(kmask);
        ++sch.n_taps;
      }
  z.sch = sch;
  z.n_total = n_total;
  z.row_elems = row;
  z.row_bytes = row_bytes;
  // resident weights when all tap chunks fit beside the slab ring
  const size_t w_bytes = static_cast<size_t>(sch.n_taps) * n_total * row_bytes;
  z.resident = (w_bytes <= 80 * 1024 && (sch.n_taps == 9 || sch.n_taps == 15)) ||
               (row_bytes < 128 && w_bytes <= 104 * 1024 && sch.n_taps == 25);
  SVS_CUDA_TRY(cudaMalloc(&z.d_weights, w_bytes));
  ZcPackArgs a{};
  a.sch = sch; a.row_elems = row; a.ct = ct; a.in_coff = g.in_coff; a.cin = g.cin; a.cout = g.cout;
  a.transposed = g.transposed ? 1 : 0; a.n_total = n_total;
  const size_t total = static_cast<size_t>(n_total) * sch.n_taps * row;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256 > 2368 ? 2368 : (total + 255) / 256);
  if (plan->w_fold[li]) {                              // a training plan has no weights yet: zc_repack_layer() per step
    if (tf32) zc_pack_weights_kernel<float><<<blocks, 256, 0, st>>>(plan->w_fold[li], a, static_cast<float*>(z.d_weights));
    else zc_pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(plan->w_fold[li], a, static_cast<__nv_bfloat16*>(z.d_weights));
    SVS_CHECK_LAUNCH("zc_pack_weights_kernel");
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(sch.n_taps) * row, static_cast<cuuint64_t>(n_total)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(sch.n_taps) * row_bytes};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(row), static_cast<cuuint32_t>(n_total)};
  int rc = encode_tensor_map(&z.tmap_b, tf32, 2, z.d_weights, dims, strides, box, row_bytes);
  if (rc != SVS_OK) return rc;
  const cuuint32_t box_half[2] = {static_cast<cuuint32_t>(row), static_cast<cuuint32_t>(n_total / 2)};
  rc = encode_tensor_map(&z.tmap_b_half, tf32, 2, z.d_weights, dims, strides, box_half, row_bytes);
  if (rc != SVS_OK) return rc;
  const cuuint32_t box_q[2] = {static_cast<cuuint32_t>(row), static_cast<cuuint32_t>(n_total / 4)};
  rc = encode_tensor_map(&z.tmap_b_quarter, tf32, 2, z.d_weights, dims, strides, box_q, row_bytes);
  if (rc != SVS_OK) return rc;
  z.enabled = true;
  return SVS_OK;
}

// (Re)packs layer li's weights from w_fold [25][cin][cout] (stream ordered; the training step calls it every iteration)
int zc_repack_layer(svs_unet_plan* plan, int li, const float* w_fold, cudaStream_t st) {
  const LayerGeom& g = kLayers[li];
  ZcLayer& z = plan->zc[li];
  if (!z.enabled) return fail(SVS_ERR_INVALID_ARG, "zc_repack_layer: layer has no zero-copy plan");
  const bool tf32 = plan->precision == SVS_PRECISION_TF32;
  ZcPackArgs a{};
  a.sch = z.sch; a.row_elems = z.row_elems; a.ct = kBufGeom[g.in_buf].c; a.in_coff = g.in_coff; a.cin = g.cin;
  a.cout = g.cout; a.transposed = g.transposed ? 1 : 0; a.n_total = z.n_total;
  const size_t total = static_cast<size_t>(z.n_total) * z.sch.n_taps * z.row_elems;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256 > 2368 ? 2368 : (total + 255) / 256);
  if (tf32) zc_pack_weights_kernel<float><<<blocks, 256, 0, st>>>(w_fold, a, static_cast<float*>(z.d_weights));
  else zc_pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w_fold, a, static_cast<__nv_bfloat16*>(z.d_weights));
  SVS_CHECK_LAUNCH("zc_pack_weights_kernel");
  return SVS_OK;
}

void zc_free_layers(svs_unet_plan* plan) {
  for (int li = 0; li < 12; ++li) {
    if (plan->zc[li].d_weights) cudaFree(plan->zc[li].d_weights);
    plan->zc[li].d_weights = nullptr;
    plan->zc[li].enabled = false;
  }
}

template <typename OutT, bool kTf32, int kBlockN, int kASlots, int kBSlots, typename Taps, int kRow = 128, int kCluster = 1,
          bool kTrim = false, bool kDual = false, int kStoreUnits = 0, bool kStoreAlias = false>
static int zc_launch_t(const CUtensorMap& ta, const CUtensorMap& tb, const ZcParams& p, cudaStream_t st,
                       const CUtensorMap* to = nullptr) {
  auto kern = zc_conv_kernel<OutT, kTf32, kBlockN, kASlots, kBSlots, Taps, kRow, kCluster, kTrim, kDual, kStoreUnits, kStoreAlias>;
  constexpr int kTiles = kDual ? 2 : 1;
  constexpr size_t smem = zc_smem_bytes<kBlockN, kTiles * kASlots, kBSlots, kRow,
                                        kStoreAlias ? 0 : kStoreUnits * 128 * 32 * static_cast<int>(sizeof(OutT))>();
  static_assert(smem <= 227 * 1024, "shared memory budget");
  if (kStoreUnits > 0 && to == nullptr) return fail(SVS_ERR_INVALID_ARG, "zc_launch_t: output tensor map missing");
  const CUtensorMap& tmo = to ? *to : ta;            // unused by the plain epilogue
  SVS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  constexpr int kAccCols = kBlockN < 32 ? 32 : kBlockN;
  int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  if (per_sm > 512 / (2 * kTiles * kAccCols)) per_sm = 512 / (2 * kTiles * kAccCols);
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = num_sms() * per_sm;
  if (grid > p.m_tiles / kTiles) grid = p.m_tiles / kTiles;
  if constexpr (kCluster == 1) {
    SVS_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(kZcThreads), smem, st, ta, tb, tmo, p));
  } else {
    grid &= ~1;                                      // whole CTA pairs; the caller guarantees an even tile count
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kZcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    SVS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, tb, tmo, p));
  }
  return SVS_OK;
}

int zc_launch_layer(const svs_unet_plan* plan, int li, const Workspace& ws, int batch, cudaStream_t st) {
  const LayerGeom& g = kLayers[li];
  ZcIo io;
  io.in = ws.buf[g.in_buf]; io.out = ws.buf[g.out_buf];
  io.out_pitch = kBufGeom[g.out_buf].c; io.out_coff = g.out_coff;
  io.bias = plan->b_fold[li]; io.act = g.act;
  return zc_launch_layer_io(plan, li, io, batch, st);
}

// The same launch on explicit buffers: the training forward reads the fp32 concat buffer of the layer's input and
// writes the raw sums (bias, no activation, no TF32 rounding) into the dense pre-BatchNorm buffer z.
int zc_launch_layer_io(const svs_unet_plan* plan, int li, const ZcIo& io, int batch, cudaStream_t st) {
  const ZcLayer& z = plan->zc[li];
  const LayerGeom& g = kLayers[li];
  const bool tf32 = plan->precision == SVS_PRECISION_TF32;
  const int es = plan->elem_size;
  CUtensorMap ta;
  {
    const cuuint64_t ct = kBufGeom[g.in_buf].c, H = g.hin, W = g.win;
    cuuint64_t dims[5], strides[4];
    if (!g.transposed) {
      dims[0] = 2 * ct; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = batch;
      strides[0] = 2 * ct * es; strides[1] = W * ct * es; strides[2] = 2 * W * ct * es; strides[3] = H * W * ct * es;
    } else {
      dims[0] = ct; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = batch;
      strides[0] = ct * es; strides[1] = W * ct * es; strides[2] = W * ct * es; strides[3] = H * W * ct * es;
    }
    const cuuint32_t box[5] = {static_cast<cuuint32_t>(z.row_elems), kZcPw, 1, kZcPh, 1};
    // narrow rows: do not let the L2 promote the request to the whole 128- / 256-byte pixel neighbourhood
    int rc = encode_tensor_map(&ta, tf32, 5, const_cast<void*>(io.in), dims, strides, box, z.row_bytes,
                               z.row_bytes < 128 ? (z.row_bytes >= 64 ? 64 : 0) : 256);
    if (rc != SVS_OK) return rc;
  }
  ZcParams p{};
  p.sch = z.sch;
  p.row_elems = z.row_elems;
  const int gh = g.transposed ? g.hin : g.hout, gw = g.transposed ? g.win : g.wout;
  p.ntw = gw / kZcBw; p.nth = gh / kZcBh;
  p.m_tiles = p.ntw * p.nth * batch;
  p.batch = batch;
  p.out = io.out;
  p.out_pitch = io.out_pitch; p.out_coff = io.out_coff;
  p.hout = g.hout; p.wout = g.wout;
  p.out_scale = g.transposed ? 2 : 1;
  p.bias = io.bias;
  p.cout_phase = g.cout;
  p.merged = g.transposed ? 1 : 0;
  p.act = io.act;
  p.keep_fp32 = io.keep_fp32;
  p.wait_first = io.wait_first;
  p.resident = z.resident ? 1 : 0;
  p.dbg = (g_tc_dbg_layer == li) ? g_tc_dbg : nullptr;
  const int n = z.n_total;
  // TMA-store epilogue (bf16 layers whose shared-memory budget has room for the staging ring): on by default,
  // SVS_ZC_TMASTORE=0 keeps the per-thread stores
  static const bool tma_store = [] { const char* e = std::getenv("SVS_ZC_TMASTORE"); return !(e && e[0] == '0'); }();
  static const bool experiments = [] {            // the opt-in variants below keep the plain epilogue
    for (const char* name : {"SVS_ZC_MCAST", "SVS_ZC_TRIM", "SVS_ZC_RING", "SVS_ZC_DUAL"}) {
      const char* e = std::getenv(name);
      if (e && e[0] != '0') return true;
    }
    return false;
  }();
  if (tma_store && !tf32 && !experiments && (li == 2 || li == 3 || li == 8 || li == 10)) {
    // 5-D output view (channel, px, x, py, batch * grid rows + y); conv layers: px = py = 1
    CUtensorMap to;
    const cuuint64_t pitch = io.out_pitch, sc = g.transposed ? 2 : 1;
    const cuuint64_t dims[5] = {pitch, sc, static_cast<cuuint64_t>(gw), sc, static_cast<cuuint64_t>(gh) * batch};
    const cuuint64_t strides[4] = {pitch * es, sc * pitch * es, static_cast<cuuint64_t>(g.wout) * pitch * es,
                                   sc * g.wout * pitch * es};
    const int box_cols = g.cout < 32 ? g.cout : 32;
    const cuuint32_t box[5] = {static_cast<cuuint32_t>(box_cols), 1, kZcBw, 1, kZcBh};
    int rc = encode_tensor_map(&to, tf32, 5, io.out, dims, strides, box, box_cols * es, 0);
    if (rc != SVS_OK) return rc;
#define SVS_ZC_STORE(LI, N, AS, BS, RES, TAPS, ROW, UNITS)                                              \
    if (li == LI && n == N && z.resident == RES && z.row_bytes == ROW && z.sch.n_slabs == TAPS::kSlabs && \
        (!RES || z.sch.n_taps == BS))                                                                     \
      return zc_launch_t<__nv_bfloat16, false, N, AS, BS, TAPS, ROW, 1, false, false, UNITS>(ta, z.tmap_b, p, st, &to);
    // one tile per CTA (batch <= 74): full-depth operand rings, the store ring aliases the slab slots; larger
    // batches keep the plain epilogue for these two layers (a dedicated ring would cost a ring slot each, measured
    // slower: conv4 13.5 vs 12.8 us, deconv3 21.0 vs 19.8 us at batch 64)
    if (p.m_tiles <= num_sms()) {
      if (li == 3 && n == 128 && !z.resident && z.sch.n_slabs == 4)
        return zc_launch_t<__nv_bfloat16, false, 128, 5, 6, ZcConvParityTaps<0xF>, 128, 1, false, false, 4, true>(ta, z.tmap_b, p, st, &to);
      if (li == 8 && n == 256 && !z.resident && z.sch.n_slabs == 4)
        return zc_launch_t<__nv_bfloat16, false, 256, 3, 4, ZcDeconvTaps<4>, 128, 1, false, false, 4, true>(ta, z.tmap_b, p, st, &to);
    }
    SVS_ZC_STORE(2, 64, 6, 25, true, ZcConvParityTaps<0x3>, 64, 2)       // conv3 (narrow rows, resident weights)
    SVS_ZC_STORE(10, 64, 4, 9, true, ZcDeconvTaps<1>, 128, 4)            // deconv5: 96 + 72 + 32 KB
#undef SVS_ZC_STORE
  }
  // bf16 layers of the reference network get compile-time tap tables; anything else (TF32 rows are 32
  // channels wide, so the slab/tap sets differ) runs the same kernel with the runtime schedule
#define SVS_ZC_NARROW(TF, LI, N, AS, BS, TAPS, ROW)                                                     \
  if (tf32 == TF && li == LI && n == N && z.resident && z.row_bytes == ROW && z.sch.n_slabs == TAPS::kSlabs && \
      z.sch.n_taps == BS) {                                                                                \
    if constexpr (TF) return zc_launch_t<float, true, N, AS, BS, TAPS, ROW>(ta, z.tmap_b, p, st);          \
    else return zc_launch_t<__nv_bfloat16, false, N, AS, BS, TAPS, ROW>(ta, z.tmap_b, p, st);              \
  }
  // narrow slab rows (the layer's own channel window), all 25 tap chunks resident
  SVS_ZC_NARROW(false, 1, 32, 8, 25, ZcConvParityTaps<0x1>, 32)      // conv2: 16 bf16 channels = 32-byte rows, 2 CTAs / SM
  SVS_ZC_NARROW(false, 2, 64, 6, 25, ZcConvParityTaps<0x3>, 64)      // conv3: 32 bf16 channels = 64-byte rows, 100 KB of weights
  SVS_ZC_NARROW(true, 1, 32, 4, 25, ZcConvParityTaps<0x3>, 64)       // conv2 TF32: 16 fp32 channels = 64-byte rows
#undef SVS_ZC_NARROW
  if (z.row_bytes != 128) return fail(SVS_ERR_NOT_IMPLEMENTED, "zc_launch_layer: no narrow-row instantiation");
  // streamed weights shared by CTA pairs (multicast halves): OFF by default.  Measured at batch 64 / 512 (bf16:
  // 225.7 vs 224.3 us, TF32: 458 vs 427 us per 64 patches): the pair's lock step costs as much as the halved weight
  // requests save — the kernels are bound by bytes INTO an SM's shared memory (~58 B/clk), which a multicast does not
  // reduce; only splitting B across the pair (cta_group::2) or two M tiles per weight pass would.
  static const bool mcast_on = [] { const char* e = std::getenv("SVS_ZC_MCAST"); return e && e[0] == '1'; }();
  const bool pair = mcast_on && !z.resident && p.m_tiles % 2 == 0 && p.m_tiles >= 2;
  // phase-trimmed merged deconvs (streamed weights): see zc_conv_kernel.  OFF by default — measured at batch 64: bf16
  // 222.8 vs 221.6 us, TF32 487.7 vs 439.3 us per forward.  Trimming removes weight bytes and MMA columns in the same
  // proportion, and the streamed-weight ring is LATENCY bound (bytes in flight per SM are capped by shared memory:
  // ring bytes / L2 latency ~ 55 B/clk), so the ratio that matters — MMA work per streamed byte — does not improve,
  // while the quarter-chunk loads and N = Cout instructions add requests.
  static const bool trim_on = [] { const char* e = std::getenv("SVS_ZC_TRIM"); return e && e[0] == '1'; }();
  const bool trim = trim_on && g.transposed && !z.resident && !pair && n % 64 == 0;
#define SVS_ZC_TRIMMED(TF, LI, N, AS, BS, TAPS)                                                      \
  if (trim && tf32 == TF && li == LI && n == N && z.sch.n_slabs == TAPS::kSlabs) {                    \
    if constexpr (TF) return zc_launch_t<float, true, N, AS, BS, TAPS, 128, 1, true>(ta, z.tmap_b_quarter, p, st);          \
    else return zc_launch_t<__nv_bfloat16, false, N, AS, BS, TAPS, 128, 1, true>(ta, z.tmap_b_quarter, p, st);              \
  }
  SVS_ZC_TRIMMED(false, 8, 256, 3, 4, ZcDeconvTaps<4>)
  SVS_ZC_TRIMMED(false, 9, 128, 2, 3, ZcDeconvTaps<2>)
  SVS_ZC_TRIMMED(true, 8, 256, 3, 4, ZcDeconvTaps<8>)
  SVS_ZC_TRIMMED(true, 9, 128, 4, 4, ZcDeconvTaps<4>)
  SVS_ZC_TRIMMED(true, 10, 64, 3, 6, ZcDeconvTaps<2>)
#undef SVS_ZC_TRIMMED
#define SVS_ZC_STATIC(TF, LI, N, AS, BS, RES, TAPS)                                                 \
  if (tf32 == TF && li == LI && n == N && z.resident == RES && z.sch.n_slabs == TAPS::kSlabs) {       \
    if constexpr (!RES) {                                                                             \
      if (pair) {                                                                                     \
        if constexpr (TF) return zc_launch_t<float, true, N, AS, BS, TAPS, 128, 2>(ta, z.tmap_b_half, p, st);          \
        else return zc_launch_t<__nv_bfloat16, false, N, AS, BS, TAPS, 128, 2>(ta, z.tmap_b_half, p, st);              \
      }                                                                                               \
    }                                                                                                 \
    if constexpr (TF) return zc_launch_t<float, true, N, AS, BS, TAPS>(ta, z.tmap_b, p, st);          \
    else return zc_launch_t<__nv_bfloat16, false, N, AS, BS, TAPS>(ta, z.tmap_b, p, st);              \
  }
  // experiment (SVS_ZC_RING=1): deeper weight rings at the expense of slab slots — the streamed-weight ring is latency
  // bound, so more bytes in flight should raise throughput
  static const bool deep_ring = [] { const char* e = std::getenv("SVS_ZC_RING"); return e && e[0] == '1'; }();
  if (deep_ring) {
    SVS_ZC_STATIC(false, 3, 128, 3, 9, false, ZcConvParityTaps<0xF>)      // conv4: 72 KB slabs + 144 KB weights
    SVS_ZC_STATIC(false, 8, 256, 2, 5, false, ZcDeconvTaps<4>)            // deconv3: 48 KB + 160 KB
    SVS_ZC_STATIC(false, 9, 128, 3, 9, false, ZcDeconvTaps<2>)            // deconv4: one CTA / SM, 72 KB + 144 KB
  }
  // two M tiles per weight pass (kDual): halves the streamed weight bytes per MMA.  Needs an even tile count and
  // enough tile pairs to fill the GPU (one CTA per SM: the two double-buffered accumulators take all of TMEM).
  static const int dual_mode = [] { const char* e = std::getenv("SVS_ZC_DUAL"); return e ? std::atoi(e) : 0; }();
  const bool dual = dual_mode != 0 && !z.resident && !pair && p.m_tiles % 2 == 0 && p.m_tiles >= 2 * num_sms();
#define SVS_ZC_DUAL(TF, LI, N, AS, BS, TAPS)                                                        \
  if (dual && tf32 == TF && li == LI && n == N && z.sch.n_slabs == TAPS::kSlabs) {                    \
    if constexpr (TF) return zc_launch_t<float, true, N, AS, BS, TAPS, 128, 1, false, true>(ta, z.tmap_b, p, st);          \
    else return zc_launch_t<__nv_bfloat16, false, N, AS, BS, TAPS, 128, 1, false, true>(ta, z.tmap_b, p, st);              \
  }
  SVS_ZC_DUAL(false, 9, 128, 2, 7, ZcDeconvTaps<2>)           // deconv4: 92 KB of slabs + 112 KB weight ring
  SVS_ZC_DUAL(false, 3, 128, 2, 7, ZcConvParityTaps<0xF>)     // conv4 (large batches only)
  SVS_ZC_DUAL(true, 9, 128, 2, 7, ZcDeconvTaps<4>)
#undef SVS_ZC_DUAL
  SVS_ZC_STATIC(false, 1, 32, 2, 15, true, ZcConv2Taps)       // 2 CTAs / SM: one CTA's epilogue hides the other's loads
  SVS_ZC_STATIC(false, 2, 64, 3, 4, false, ZcConvParityTaps<0xC>)
  SVS_ZC_STATIC(false, 3, 128, 5, 6, false, ZcConvParityTaps<0xF>)
  SVS_ZC_STATIC(false, 8, 256, 3, 4, false, ZcDeconvTaps<4>)
  SVS_ZC_STATIC(false, 9, 128, 2, 3, false, ZcDeconvTaps<2>)
  SVS_ZC_STATIC(false, 10, 64, 4, 9, true, ZcDeconvTaps<1>)
  // TF32: 128-byte rows hold 32 channels, so every layer has twice the slabs of its bf16 form
  SVS_ZC_STATIC(true, 1, 32, 3, 8, false, ZcConvParityTaps<0xC>)
  SVS_ZC_STATIC(true, 2, 64, 3, 4, false, ZcConvParityTaps<0xF>)
  SVS_ZC_STATIC(true, 3, 128, 4, 4, false, ZcConvParity2Taps)
  SVS_ZC_STATIC(true, 8, 256, 3, 4, false, ZcDeconvTaps<8>)
  SVS_ZC_STATIC(true, 9, 128, 4, 4, false, ZcDeconvTaps<4>)
  SVS_ZC_STATIC(true, 10, 64, 3, 6, false, ZcDeconvTaps<2>)
#undef SVS_ZC_STATIC
#define SVS_ZC_CASE(N, AS, BS, RES)                                                               \
  if (n == N && z.resident == RES && (!RES || z.sch.n_taps == BS))                                 \
    return tf32 ? zc_launch_t<float, true, N, AS, BS, ZcRuntimeTaps>(ta, z.tmap_b, p, st)          \
                : zc_launch_t<__nv_bfloat16, false, N, AS, BS, ZcRuntimeTaps>(ta, z.tmap_b, p, st);
  SVS_ZC_CASE(256, 3, 4, false)
  SVS_ZC_CASE(128, 5, 6, false)
  SVS_ZC_CASE(64, 5, 8, false)
  SVS_ZC_CASE(32, 5, 8, false)
  SVS_ZC_CASE(64, 4, 9, true)      // deconv5: 9 x 8 KB resident
  SVS_ZC_CASE(32, 4, 15, true)     // conv2: 15 x 4 KB resident
#undef SVS_ZC_CASE
  return fail(SVS_ERR_NOT_IMPLEMENTED, "zc_launch_layer: no instantiation for N=" + std::to_string(n));
}

}  // namespace svs
