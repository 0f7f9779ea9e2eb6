// K6: weight gradient of the 5x5 stride-2 convolutions / transposed convolutions on tcgen05 (TF32).
//
// Reference: loss.backward() at train.py:298 through nn.Conv2d / nn.ConvTranspose2d (model.py:47-109).
//
//   conv   : dW[co][ci][kh][kw] = sum_{b,oy,ox} dz[b,oy,ox,co] * x[b, 2oy+kh-2, 2ox+kw-2, ci]
//   deconv : dW[ci][co][kh][kw] = sum_{b,iy,ix} x[b,iy,ix,ci]  * dz[b, 2iy+kh-2, 2ix+kw-2, co]
//
// Both are the same contraction: S lives on the SMALL grid (conv: dz, deconv: x), L on the 2x grid (conv: x,
// deconv: dz), and  R[tap][m][n] = sum_pixels S[p][m] * L[p + tap][n].  The reduction runs over PIXELS, and both
// tensors are NHWC, i.e. pixel-major with channels contiguous: exactly the "MN-major" operand form of tcgen05
// (K = pixels down the rows, M / N = channels along the 128-byte / 64-byte swizzled rows), so TMA lands the tiles
// as they lie in memory and no transpose is needed.
//
//   A = S tile  [128 pixels][128 channels]  fp32, four 32-channel chunks of 128-byte rows, MN-major (chunk pitch =
//       leading byte offset); channels beyond the tensor are TMA zero fill
//   B = L tile  [128 pixels][32 channels]   fp32, 128-byte rows, MN-major, ONE PER TAP: the stride-2 window is a
//       plain box of the parity-split rank-5 view (pw*C + c, W/2, ph, H/2, B); the zero padding of the convolution
//       is the TMA out-of-bounds fill
//   Both use the 128-byte swizzle with 32-BYTE atoms (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B on the TMA side,
//   UMMA layout type SWIZZLE_128B_BASE32B): it is the only shared-memory layout tcgen05 accepts for MN-major TF32
//   operands (with the ordinary 16-byte-atom swizzles the MMA silently produces zeros: measured).  A swizzle atom is
//   4 pixel rows x 128 B, so one K = 8 instruction spans two atoms (stride byte offset 512 B).
//   D = 13 accumulators [128][32] fp32 in TMEM (416 of 512 columns), one per tap.  25 taps x 32 columns do not fit,
//       so a CTA makes TWO passes over its pixel tiles: taps 0..12, then taps 12..24 (tap 12 is simply computed in
//       both).  kind::tf32, K = 8 pixels per instruction -> 2 x 13 x 16 = 416 MMAs per 128-pixel tile.
//
// With all taps of a pass resident in TMEM, S and L cross L2 -> shared memory twice per (m tile, n tile); the kernel
// is bound by MMA issue (N = 32 instructions), not by bytes.  Pixel tiles are split over CTAs; every CTA writes its
// partial R[split][tap][m][n] and wgrad_tc_finalize_kernel adds the splits in fixed order (deterministic, no atomics)
// while scattering into the torch weight layout ((m * Ln + n) * 25 + tap for both cases).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
#include "unet_internal.cuh"
#include "tc_ptx.cuh"
#include "tc_conv_common.cuh"

#include <cstdlib>

namespace svs {

constexpr int kWgThreads = 192;
constexpr int kWgASlots = 2;
constexpr int kWgAChunkBytes = 128 * 128;              // [128 pixels][32 channels] fp32
constexpr int kWgABytes = 4 * kWgAChunkBytes;
// kN = channels of L per CTA (32 or 64).  512 TMEM columns hold 512 / kN - 3 accumulators of kN columns with room
// to spare (13 for kN = 32, 7 for kN = 64), so the 25 taps take 2 or 4 passes over the CTA's pixel tiles; successive
// passes overlap by one tap (pass p covers taps p (T-1) .. p (T-1) + T-1), which keeps every pass the same length.
// The wide form needs half the MMAs per channel but re-loads S four times instead of twice; measured equal or slightly
// slower at batch 64, so N = 32 is the default and N = 64 stays as an option (SVS_WGRAD_N64=1, parity-tested).
template <int kN> struct WgCfg {
  static constexpr int kTaps = kN == 32 ? 13 : 7;
  static constexpr int kPasses = kN == 32 ? 2 : 4;
  static constexpr int kBSlots = kN == 32 ? 4 : 3;
  static constexpr int kBBytes = (kN / 32) * kWgAChunkBytes;      // kN / 32 chunks of [128 pixels][32 channels] fp32
  static constexpr size_t kSmem = static_cast<size_t>(kWgASlots) * kWgABytes + kBSlots * kBBytes + 1024 + 256;
  static_assert((kPasses - 1) * (kTaps - 1) + kTaps == 25, "passes must cover the 25 taps");
  static_assert(kTaps * kN <= 512, "TMEM columns");
};

struct WgParams {
  int m_tiles, n_tiles, splits;          // work units (m tile, n tile) x pixel-tile splits
  int pix_tiles;                         // 128-pixel tiles of the small grid
  int ntw, nth;                          // tiles along w / h of one image
  int bw, bh, nb;
  int s_c, l_c;                          // channels of S (M) and L (N)
  int l_pitch, l_coff;                   // L: channels per pixel / first channel
  int s_coff;
  float* partial;                        // [splits][25][s_c][l_c]
};

// MN-major operand descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout type 1 (SWIZZLE_128B_BASE32B) << 61.  LBO = pitch of the 32-channel chunks along M,
// SBO = pitch of the 4-row swizzle atoms along K.
__device__ __forceinline__ uint64_t wg_desc(uint32_t addr, uint32_t lbo_bytes) {
  return static_cast<uint64_t>((addr & 0x3FFFF) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}

template <int kWgN>
__global__ void __launch_bounds__(kWgThreads)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_s, const __grid_constant__ CUtensorMap tmap_l,
                const __grid_constant__ WgParams p) {
  constexpr int kWgTapsPerPass = WgCfg<kWgN>::kTaps, kWgPasses = WgCfg<kWgN>::kPasses, kWgBSlots = WgCfg<kWgN>::kBSlots;
  constexpr int kWgBBytes = WgCfg<kWgN>::kBBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_base = smem_base + kWgASlots * kWgABytes;
  const size_t bar_off = static_cast<size_t>(kWgASlots) * kWgABytes + static_cast<size_t>(kWgBSlots) * kWgBBytes;
  const uint32_t bar_base = smem_base + static_cast<uint32_t>(bar_off);
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto empty_a = [&](int s) { return bar_base + 8u * (kWgASlots + s); };
  auto full_b = [&](int s) { return bar_base + 8u * (2 * kWgASlots + s); };
  auto empty_b = [&](int s) { return bar_base + 8u * (2 * kWgASlots + kWgBSlots + s); };
  const uint32_t acc_full = bar_base + 8u * (2 * kWgASlots + 2 * kWgBSlots);
  const uint32_t acc_empty = acc_full + 8u;
  const uint32_t tmem_slot = acc_full + 16u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * (2 * kWgASlots + 2 * kWgBSlots + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA -> (unit, split): the splits of one unit are adjacent so that they share S / L tiles in L2
  const int unit = blockIdx.x / p.splits, split = blockIdx.x - unit * p.splits;
  const int mt = unit / p.n_tiles, nt = unit - mt * p.n_tiles;
  const int per = (p.pix_tiles + p.splits - 1) / p.splits;
  const int t_begin = split * per, t_end = min(p.pix_tiles, t_begin + per);
  const int n_my = max(0, t_end - t_begin);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgASlots; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
    for (int s = 0; s < kWgBSlots; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_s);
    tma_prefetch_desc(&tmap_l);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===== TMA producer =====
    int it_a = 0, it_b = 0;
    for (int pass = 0; pass < kWgPasses; ++pass) {
      const int tap0 = pass * (kWgTapsPerPass - 1);
      for (int i = 0; i < n_my; ++i, ++it_a) {
        const int tile = t_begin + i;
        const int tw = tile % p.ntw, th = (tile / p.ntw) % p.nth, tb = tile / (p.ntw * p.nth);
        const int sa = it_a % kWgASlots;
        mbar_wait(empty_a(sa), ((it_a / kWgASlots) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(full_a(sa), kWgABytes);
          // four 32-channel chunks; chunks past the tensor's channels are zero filled by TMA (and still count bytes)
          for (int c = 0; c < 4; ++c)
            tma_load_5d(smem_base + sa * kWgABytes + c * kWgAChunkBytes, &tmap_s, full_a(sa),
                        p.s_coff + mt * 128 + c * 32, tw * p.bw, 0, th * p.bh, tb * p.nb);
        }
        __syncwarp();
        for (int t = 0; t < kWgTapsPerPass; ++t, ++it_b) {
          const int tap = tap0 + t;
          const int kh = tap / 5, kw = tap - 5 * kh;
          const int qh = kh - 2, qw = kw - 2;
          const int ph = qh & 1, pw = qw & 1;
          const int dh = (qh - ph) / 2, dw = (qw - pw) / 2;
          const int sb = it_b % kWgBSlots;
          mbar_wait(empty_b(sb), ((it_b / kWgBSlots) & 1) ^ 1);
          if (elect_one_sync()) {
            mbar_expect_tx(full_b(sb), kWgBBytes);
            for (int c = 0; c < kWgN / 32; ++c)
              tma_load_5d(b_base + sb * kWgBBytes + c * kWgAChunkBytes, &tmap_l, full_b(sb),
                          pw * p.l_pitch + p.l_coff + nt * kWgN + c * 32, tw * p.bw + dw, ph, th * p.bh + dh, tb * p.nb);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: per tile 13 taps x 16 K steps; descriptors are "uniform base + compile-time constant" =====
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               (static_cast<uint32_t>(kWgN >> 3) << 17) | ((128u >> 4) << 24);
    int it_a = 0, it_b = 0;
    for (int pass = 0; pass < kWgPasses; ++pass) {
      if (pass > 0) {                                    // the epilogue has drained the accumulators of the last pass
        mbar_wait(acc_empty, (pass - 1) & 1);
        tc_fence_after();
      }
      for (int i = 0; i < n_my; ++i, ++it_a) {
        const int sa = it_a % kWgASlots;
        mbar_wait(full_a(sa), (it_a / kWgASlots) & 1);
        tc_fence_after();
        for (int t = 0; t < kWgTapsPerPass; ++t, ++it_b) {
          const int sb = it_b % kWgBSlots;
          mbar_wait(full_b(sb), (it_b / kWgBSlots) & 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + t * kWgN;
          const uint32_t acc0 = i > 0 ? 1u : 0u;
          dispatch_stage<0, kWgASlots * kWgBSlots>(sa * kWgBSlots + sb, [&](auto sc) {
            constexpr int SA = decltype(sc)::value / kWgBSlots, SB = decltype(sc)::value % kWgBSlots;
            const uint64_t da = wg_desc(smem_base + SA * kWgABytes, kWgAChunkBytes);
            const uint64_t db = wg_desc(smem_base + kWgASlots * kWgABytes + SB * kWgBBytes, kWgAChunkBytes);
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < 16; ++j)       // K step j = pixel rows 8j .. 8j+7: +1024 B on both start addresses
                umma<true>(tmem_d, da + static_cast<uint64_t>(j * (1024 >> 4)), db + static_cast<uint64_t>(j * (1024 >> 4)),
                           idesc, j > 0 ? 1u : acc0);
              umma_commit(bar_base + 8u * (2 * kWgASlots + kWgBSlots + SB));      // empty_b(SB)
            }
            __syncwarp();
          });
        }
        if (elect_one_sync()) umma_commit(empty_a(sa));
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM lane = channel m of S; 13 x 32 columns per pass -> partial[split][tap][m][n] =====
    const int q = warp & 3;
    const int r = 32 * q + lane;
    const int m = mt * 128 + r;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16);
    const size_t tap_pitch = static_cast<size_t>(p.s_c) * p.l_c;
    float* dst0 = p.partial + (static_cast<size_t>(split) * 25 * p.s_c + m) * p.l_c + nt * kWgN;
    const int n_valid = min(kWgN, p.l_c - nt * kWgN);            // 32, or 16 for the 16-channel layers
    for (int pass = 0; pass < kWgPasses; ++pass) {
      const int tap0 = pass * (kWgTapsPerPass - 1);
      mbar_wait(acc_full, pass & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < kWgTapsPerPass; ++t) {
#pragma unroll
        for (int h = 0; h < kWgN; h += 32) {
          uint32_t v[32];
          if (n_my > 0) {
            tmem_ld16(taddr + t * kWgN + h, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
            tmem_ld16(taddr + t * kWgN + h + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = 0u;
          }
          if (m < p.s_c) {
            float4* d4 = reinterpret_cast<float4*>(dst0 + (tap0 + t) * tap_pitch + h);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (h + 4 * k < n_valid)
                d4[k] = make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]), __uint_as_float(v[4 * k + 2]),
                                    __uint_as_float(v[4 * k + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// GROUPED form (default).  In wgrad_tc_kernel every tap is its own N = 32 instruction, and each of them re-reads the
// whole A tile (128 channels x 8 pixels = 4 KB) from shared memory for 1 KB of B: the tensor core's operand fetch,
// not its arithmetic, sets the pace (measured 41 cycles per instruction against a 16-cycle MMA floor; 14 % tensor-pipe
// activity in ncu).  In the MN-major layout the N dimension of B is a sequence of 32-channel chunks at a constant
// pitch (the descriptor's leading byte offset) -- exactly what A's four chunks already are -- so the B tiles of
// EIGHT TAPS, landed in eight consecutive 8 KB chunks, are ONE N = 256 operand: one instruction per K step multiplies
// the A tile with eight taps at once and writes accumulator columns [32 tap, 32 tap + 32) for all of them.  Per pass
// the 13 taps are two groups (8 taps: N = 256, 5 taps: N = 160): 16 instructions per pixel tile instead of 208, and
// 12 KB of operand fetch per eight taps instead of 40 KB.  Pixel tiles are 64 pixels so that two A tiles (2 x 32 KB)
// and two tap groups (2 x 64 KB) fit in shared memory.
constexpr int kGPix = 64;
constexpr int kGChunk = kGPix * 128;                  // [64 pixels][32 channels] fp32
constexpr int kGABytes = 4 * kGChunk;
constexpr int kGASlots = 2;
constexpr int kGGroupTaps = 8;
constexpr int kGBSlotBytes = kGGroupTaps * kGChunk;
constexpr int kGBSlots = 2;
constexpr size_t kGSmem = static_cast<size_t>(kGASlots) * kGABytes + static_cast<size_t>(kGBSlots) * kGBSlotBytes + 1024 + 256;
static_assert(kGSmem <= 227 * 1024, "shared memory budget");

__global__ void __launch_bounds__(kWgThreads)
wgrad_tc_grouped_kernel(const __grid_constant__ CUtensorMap tmap_s, const __grid_constant__ CUtensorMap tmap_l,
                        const __grid_constant__ WgParams p) {
  constexpr int kTaps = 13, kPasses = 2, kN = 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_base = smem_base + kGASlots * kGABytes;
  const size_t bar_off = static_cast<size_t>(kGASlots) * kGABytes + static_cast<size_t>(kGBSlots) * kGBSlotBytes;
  const uint32_t bar_base = smem_base + static_cast<uint32_t>(bar_off);
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto empty_a = [&](int s) { return bar_base + 8u * (kGASlots + s); };
  auto full_b = [&](int s) { return bar_base + 8u * (2 * kGASlots + s); };
  auto empty_b = [&](int s) { return bar_base + 8u * (2 * kGASlots + kGBSlots + s); };
  const uint32_t acc_full = bar_base + 8u * (2 * kGASlots + 2 * kGBSlots);
  const uint32_t acc_empty = acc_full + 8u;
  const uint32_t tmem_slot = acc_full + 16u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * (2 * kGASlots + 2 * kGBSlots + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x / p.splits, split = blockIdx.x - unit * p.splits;
  const int mt = unit / p.n_tiles, nt = unit - mt * p.n_tiles;
  const int per = (p.pix_tiles + p.splits - 1) / p.splits;
  const int t_begin = split * per, t_end = min(p.pix_tiles, t_begin + per);
  const int n_my = max(0, t_end - t_begin);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGASlots; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
    for (int s = 0; s < kGBSlots; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_s);
    tma_prefetch_desc(&tmap_l);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===== TMA producer: per pixel tile the A tile, then the two tap groups of the pass =====
    int it_a = 0, it_b = 0;
    for (int pass = 0; pass < kPasses; ++pass) {
      const int tap0 = pass * (kTaps - 1);
      for (int i = 0; i < n_my; ++i, ++it_a) {
        const int tile = t_begin + i;
        const int tw = tile % p.ntw, th = (tile / p.ntw) % p.nth, tb = tile / (p.ntw * p.nth);
        const int sa = it_a % kGASlots;
        mbar_wait(empty_a(sa), ((it_a / kGASlots) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(full_a(sa), kGABytes);
          for (int c = 0; c < 4; ++c)
            tma_load_5d(smem_base + sa * kGABytes + c * kGChunk, &tmap_s, full_a(sa), p.s_coff + mt * 128 + c * 32,
                        tw * p.bw, 0, th * p.bh, tb * p.nb);
        }
        __syncwarp();
        for (int g = 0; g < 2; ++g, ++it_b) {
          const int first = g * kGGroupTaps, n_t = g == 0 ? kGGroupTaps : kTaps - kGGroupTaps;
          const int sb = it_b % kGBSlots;
          mbar_wait(empty_b(sb), ((it_b / kGBSlots) & 1) ^ 1);
          if (elect_one_sync()) {
            mbar_expect_tx(full_b(sb), static_cast<uint32_t>(n_t) * kGChunk);
            for (int t = 0; t < n_t; ++t) {
              const int tap = tap0 + first + t;
              const int kh = tap / 5, kw = tap - 5 * kh;
              const int qh = kh - 2, qw = kw - 2;
              const int ph = qh & 1, pw = qw & 1;
              const int dh = (qh - ph) / 2, dw = (qw - pw) / 2;
              tma_load_5d(b_base + sb * kGBSlotBytes + t * kGChunk, &tmap_l, full_b(sb),
                          pw * p.l_pitch + p.l_coff + nt * kN, tw * p.bw + dw, ph, th * p.bh + dh, tb * p.nb);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: per pixel tile 2 groups x 8 K steps =====
    constexpr uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((128u >> 4) << 24);
    constexpr uint32_t idesc_g0 = idesc_base | (static_cast<uint32_t>((kGGroupTaps * kN) >> 3) << 17);            // N = 256
    constexpr uint32_t idesc_g1 = idesc_base | (static_cast<uint32_t>(((kTaps - kGGroupTaps) * kN) >> 3) << 17);  // N = 160
    int it_a = 0, it_b = 0;
    for (int pass = 0; pass < kPasses; ++pass) {
      if (pass > 0) {                                    // the epilogue has drained the accumulators of the last pass
        mbar_wait(acc_empty, (pass - 1) & 1);
        tc_fence_after();
      }
      for (int i = 0; i < n_my; ++i, ++it_a) {
        const int sa = it_a % kGASlots;
        mbar_wait(full_a(sa), (it_a / kGASlots) & 1);
        tc_fence_after();
        const uint32_t acc0 = i > 0 ? 1u : 0u;
        for (int g = 0; g < 2; ++g, ++it_b) {
          const int sb = it_b % kGBSlots;
          mbar_wait(full_b(sb), (it_b / kGBSlots) & 1);
          tc_fence_after();
          dispatch_stage<0, 8>(sa * 4 + sb * 2 + g, [&](auto sc) {
            constexpr int SA = decltype(sc)::value >> 2, SB = (decltype(sc)::value >> 1) & 1, G = decltype(sc)::value & 1;
            const uint64_t da = wg_desc(smem_base + SA * kGABytes, kGChunk);
            const uint64_t db = wg_desc(smem_base + kGASlots * kGABytes + SB * kGBSlotBytes, kGChunk);
            const uint32_t tmem_d = tmem_base + G * kGGroupTaps * kN;
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < kGPix / 8; ++j)      // K step j = pixel rows 8j .. 8j+7: +1024 B on both start addresses
                umma<true>(tmem_d, da + static_cast<uint64_t>(j * (1024 >> 4)), db + static_cast<uint64_t>(j * (1024 >> 4)),
                           G == 0 ? idesc_g0 : idesc_g1, j > 0 ? 1u : acc0);
              umma_commit(bar_base + 8u * (2 * kGASlots + kGBSlots + SB));      // empty_b(SB)
            }
            __syncwarp();
          });
        }
        if (elect_one_sync()) umma_commit(empty_a(sa));
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM lane = channel m of S; 13 x 32 columns per pass -> partial[split][tap][m][n] =====
    const int q = warp & 3;
    const int r = 32 * q + lane;
    const int m = mt * 128 + r;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16);
    const size_t tap_pitch = static_cast<size_t>(p.s_c) * p.l_c;
    float* dst0 = p.partial + (static_cast<size_t>(split) * 25 * p.s_c + m) * p.l_c + nt * kN;
    const int n_valid = min(kN, p.l_c - nt * kN);                // 32, or 16 for the 16-channel layers
    for (int pass = 0; pass < kPasses; ++pass) {
      const int tap0 = pass * (kTaps - 1);
      mbar_wait(acc_full, pass & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < kTaps; ++t) {
        uint32_t v[32];
        if (n_my > 0) {
          tmem_ld16(taddr + t * kN, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          tmem_ld16(taddr + t * kN + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = 0u;
        }
        if (m < p.s_c) {
          float4* d4 = reinterpret_cast<float4*>(dst0 + (tap0 + t) * tap_pitch);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (4 * k < n_valid)
              d4[k] = make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]), __uint_as_float(v[4 * k + 2]),
                                  __uint_as_float(v[4 * k + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// grad_w[(m * l_c + n) * 25 + tap] = sum over splits (fixed order) of partial[split][tap][m][n]
__global__ void __launch_bounds__(256)
wgrad_tc_finalize_kernel(const float* __restrict__ partial, int splits, int s_c, int l_c, float* __restrict__ grad_w) {
  const int total = 25 * s_c * l_c;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    // four independent chains keep four loads in flight; the combination order is fixed (deterministic)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 4 <= splits; k += 4) {
      s0 += partial[static_cast<size_t>(k) * total + i];
      s1 += partial[static_cast<size_t>(k + 1) * total + i];
      s2 += partial[static_cast<size_t>(k + 2) * total + i];
      s3 += partial[static_cast<size_t>(k + 3) * total + i];
    }
    for (; k < splits; ++k) s0 += partial[static_cast<size_t>(k) * total + i];
    const float s = (s0 + s1) + (s2 + s3);
    const int n = i % l_c, m = (i / l_c) % s_c, tap = i / (l_c * s_c);
    grad_w[(static_cast<size_t>(m) * l_c + n) * 25 + tap] = s;
  }
}

// The same sum for layers with MANY splits and few outputs (the shallow layers: one or two work units, 74-148 splits).
// A CTA finishes 32 consecutive outputs: thread (group g of 8, output o) adds the splits k = g, g + 8, ... and the
// eight group sums are combined in shared memory in fixed order (deterministic, no atomics).  One thread per output
// walking all ~148 splits left a 50-CTA grid waiting on ~37 dependent L2 round trips (23-30 us per layer).
constexpr int kWgFinGroups = 8;
__global__ void __launch_bounds__(256)
wgrad_tc_finalize_many_kernel(const float* __restrict__ partial, int splits, int s_c, int l_c, float* __restrict__ grad_w) {
  __shared__ float red[kWgFinGroups][32];
  const int total = 25 * s_c * l_c;
  const int o = threadIdx.x & 31, g = threadIdx.x >> 5;
  for (int base = blockIdx.x * 32; base < total; base += gridDim.x * 32) {
    const int i = base + o;
    float s0 = 0.f, s1 = 0.f;
    if (i < total) {
      int k = g;
      for (; k + kWgFinGroups < splits; k += 2 * kWgFinGroups) {
        s0 += partial[static_cast<size_t>(k) * total + i];
        s1 += partial[static_cast<size_t>(k + kWgFinGroups) * total + i];
      }
      if (k < splits) s0 += partial[static_cast<size_t>(k) * total + i];
    }
    red[g][o] = s0 + s1;
    __syncthreads();
    if (g == 0 && i < total) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < kWgFinGroups; ++q) s += red[q][o];
      const int n = i % l_c, m = (i / l_c) % s_c, tap = i / (l_c * s_c);
      grad_w[(static_cast<size_t>(m) * l_c + n) * 25 + tap] = s;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// host
struct WgGeom { int bw, bh, nb, ntw, nth, pix_tiles, m_tiles, n_tiles, splits; };

static int wg_n(int l_c) {
  // N = 64 (four passes) is parity-green but not faster: the kernel is bound by operand bytes into shared memory, and
  // four passes re-load S twice as often as two (measured at batch 64: 4.0 vs 3.9 ms per training step) -> opt-in
  static const bool wide = [] { const char* e = std::getenv("SVS_WGRAD_N64"); return e && e[0] == '1'; }();
  return (l_c % 64 == 0 && wide) ? 64 : 32;
}

// SVS_WGRAD_GROUPED=0 selects the one-instruction-per-tap kernel (wgrad_tc_kernel)
static bool wg_grouped(int l_c) {
  static const bool on = [] { const char* e = std::getenv("SVS_WGRAD_GROUPED"); return !(e && e[0] == '0'); }();
  return on && wg_n(l_c) == 32;
}

static WgGeom wg_geom(int gh, int gw, int batch, int s_c, int l_c) {
  WgGeom g{};
  const int kWgN = wg_n(l_c);
  const int pix = wg_grouped(l_c) ? kGPix : 128;                 // pixels per tile
  g.bw = gw < 16 ? gw : 16;
  g.bh = gh < pix / g.bw ? gh : pix / g.bw;
  g.nb = pix / (g.bw * g.bh);
  g.ntw = gw / g.bw; g.nth = gh / g.bh;
  g.pix_tiles = g.ntw * g.nth * ((batch + g.nb - 1) / g.nb);
  g.m_tiles = (s_c + 127) / 128;
  g.n_tiles = (l_c + kWgN - 1) / kWgN;
  const int units = g.m_tiles * g.n_tiles;
  int s = (num_sms() + units - 1) / units;                       // about one CTA per SM
  if (s > g.pix_tiles) s = g.pix_tiles;
  if (s < 1) s = 1;
  g.splits = s;
  return g;
}

bool wgrad_tc_supported(int gh, int gw, int s_c, int l_c, int s_pitch, int l_pitch, int s_coff, int l_coff) {
  static const bool off = [] { const char* e = std::getenv("SVS_WGRAD_TC_DISABLE"); return e && e[0] == '1'; }();
  if (off) return false;
  const int bw = gw < 16 ? gw : 16;
  if (kGPix % bw != 0) return false;                             // both tile sizes (64 / 128 pixels) must decompose
  for (int pix : {kGPix, 128}) {
    const int bh = gh < pix / bw ? gh : pix / bw;
    if (pix % (bw * bh) != 0 || gw % bw != 0 || gh % bh != 0) return false;
  }
  return s_c % 32 == 0 && l_c % 16 == 0 && s_pitch % 4 == 0 && l_pitch % 4 == 0 && s_coff % 4 == 0 && l_coff % 4 == 0;
}

size_t wgrad_tc_partial_floats(int gh, int gw, int batch, int s_c, int l_c) {
  const WgGeom g = wg_geom(gh, gw, batch, s_c, l_c);
  return static_cast<size_t>(g.splits) * 25 * s_c * l_c;
}

// S: fp32 NHWC [batch][gh][gw][s_pitch], channels [s_coff, s_coff + s_c);  L: [batch][2gh][2gw][l_pitch], channels
// [l_coff, l_coff + l_c).  Both must hold TF32-representable values (the tensor core truncates).  grad_w: torch layout
// (s_c, l_c, 5, 5), overwritten.
int wgrad_tc_launch(const float* S, int s_pitch, int s_coff, int s_c, const float* L, int l_pitch, int l_coff, int l_c,
                    int gh, int gw, int batch, float* partial, size_t partial_floats, float* grad_w, cudaStream_t st) {
  const WgGeom g = wg_geom(gh, gw, batch, s_c, l_c);
  if (partial_floats < static_cast<size_t>(g.splits) * 25 * s_c * l_c)
    return fail(SVS_ERR_WORKSPACE, "wgrad_tc_launch: partial buffer too small");
  CUtensorMap ts, tl;
  {
    // S: (c, W, 1, H, B), box (32, bw, 1, bh, nb), 128-byte rows
    const cuuint64_t dims[5] = {static_cast<cuuint64_t>(s_pitch), static_cast<cuuint64_t>(gw), 1,
                                static_cast<cuuint64_t>(gh), static_cast<cuuint64_t>(batch)};
    const cuuint64_t strides[4] = {static_cast<cuuint64_t>(s_pitch) * 4, static_cast<cuuint64_t>(gw) * s_pitch * 4,
                                   static_cast<cuuint64_t>(gw) * s_pitch * 4,
                                   static_cast<cuuint64_t>(gh) * gw * s_pitch * 4};
    const cuuint32_t box[5] = {32, static_cast<cuuint32_t>(g.bw), 1, static_cast<cuuint32_t>(g.bh),
                               static_cast<cuuint32_t>(g.nb)};
    int rc = encode_tensor_map(&ts, true, 5, const_cast<float*>(S), dims, strides, box, 12832);
    if (rc != SVS_OK) return rc;
  }
  {
    // L: parity-split view (pw * C + c, W/2, ph, H/2, B) of the 2x grid, box (32, bw, 1, bh, nb), 128-byte rows
    const cuuint64_t H = 2 * gh, W = 2 * gw, ct = l_pitch;
    const cuuint64_t dims[5] = {2 * ct, W / 2, 2, H / 2, static_cast<cuuint64_t>(batch)};
    const cuuint64_t strides[4] = {2 * ct * 4, W * ct * 4, 2 * W * ct * 4, H * W * ct * 4};
    const cuuint32_t box[5] = {32, static_cast<cuuint32_t>(g.bw), 1, static_cast<cuuint32_t>(g.bh),
                               static_cast<cuuint32_t>(g.nb)};
    int rc = encode_tensor_map(&tl, true, 5, const_cast<float*>(L), dims, strides, box, 12832);
    if (rc != SVS_OK) return rc;
  }
  WgParams p{};
  p.m_tiles = g.m_tiles; p.n_tiles = g.n_tiles; p.splits = g.splits;
  p.pix_tiles = g.pix_tiles; p.ntw = g.ntw; p.nth = g.nth;
  p.bw = g.bw; p.bh = g.bh; p.nb = g.nb;
  p.s_c = s_c; p.l_c = l_c; p.l_pitch = l_pitch; p.l_coff = l_coff; p.s_coff = s_coff;
  p.partial = partial;
  const int grid = g.m_tiles * g.n_tiles * g.splits;
  if (wg_grouped(l_c)) {
    SVS_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(kGSmem)));
    wgrad_tc_grouped_kernel<<<grid, kWgThreads, kGSmem, st>>>(ts, tl, p);
  } else if (wg_n(l_c) == 64) {
    SVS_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(WgCfg<64>::kSmem)));
    wgrad_tc_kernel<64><<<grid, kWgThreads, WgCfg<64>::kSmem, st>>>(ts, tl, p);
  } else {
    SVS_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(WgCfg<32>::kSmem)));
    wgrad_tc_kernel<32><<<grid, kWgThreads, WgCfg<32>::kSmem, st>>>(ts, tl, p);
  }
  SVS_CHECK_LAUNCH("wgrad_tc_kernel");
  const int total = 25 * s_c * l_c;
  if (g.splits >= 32) {                              // many splits, few outputs: 8 split groups per output
    int blocks = (total + 31) / 32;
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_tc_finalize_many_kernel<<<blocks, 256, 0, st>>>(partial, g.splits, s_c, l_c, grad_w);
  } else {
    int blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_tc_finalize_kernel<<<blocks, 256, 0, st>>>(partial, g.splits, s_c, l_c, grad_w);
  }
  SVS_CHECK_LAUNCH("wgrad_tc_finalize_kernel");
  return SVS_OK;
}

}  // namespace svs

extern "C" size_t svs_conv_wgrad_partial_floats(int gh, int gw, int batch, int s_c, int l_c) {
  if (gh <= 0 || gw <= 0 || batch <= 0 || s_c <= 0 || l_c <= 0) return 0;
  return svs::wgrad_tc_partial_floats(gh, gw, batch, s_c, l_c);
}

extern "C" int svs_conv_wgrad_tf32(const float* small, int s_pitch, int s_coff, int s_c, const float* large,
                                   int l_pitch, int l_coff, int l_c, int gh, int gw, int batch, float* partial,
                                   size_t partial_floats, float* grad_w, void* stream) {
  using namespace svs;
  SVS_REQUIRE(small && large && partial && grad_w, "svs_conv_wgrad_tf32: null pointer");
  SVS_REQUIRE(batch > 0 && gh > 0 && gw > 0, "svs_conv_wgrad_tf32: bad sizes");
  SVS_REQUIRE(s_coff >= 0 && l_coff >= 0 && s_coff + s_c <= s_pitch && l_coff + l_c <= l_pitch,
              "svs_conv_wgrad_tf32: channel window outside the buffer");
  if (!wgrad_tc_supported(gh, gw, s_c, l_c, s_pitch, l_pitch, s_coff, l_coff))
    return fail(SVS_ERR_INVALID_ARG, "svs_conv_wgrad_tf32: unsupported shape (s_c % 32, l_c % 16, power-of-two grid)");
  SVS_REQUIRE((reinterpret_cast<uintptr_t>(small) & 15) == 0 && (reinterpret_cast<uintptr_t>(large) & 15) == 0,
              "svs_conv_wgrad_tf32: operands must be 16-byte aligned");
  int dev = 0;
  SVS_CUDA_TRY(cudaGetDevice(&dev));
  int rc = svs_device_check(dev);
  if (rc != SVS_OK) return rc;
  return wgrad_tc_launch(small, s_pitch, s_coff, s_c, large, l_pitch, l_coff, l_c, gh, gw, batch, partial,
                         partial_floats, grad_w, static_cast<cudaStream_t>(stream));
}
