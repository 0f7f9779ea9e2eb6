// K3/K4: 5x5 stride-2 convolution and transposed convolution as implicit GEMM on the sm_100a
// tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// GEMM view (reference model.py:52-108):   D[M = pixels][N = Cout] = A[M][K] * B[N][K]^T
//   conv   : M runs over OUTPUT pixels, K = 25 taps x Cin.  The NHWC input [B][H][W][Ct] is viewed
//            through a rank-5 tensor map  (pw*Ct + c, W/2, ph, H/2, B)  — rows and columns split
//            by parity — so that tap (kh, kw) of a stride-2 window is a plain box at offset
//            (dw, ph, dh) = (floor((kw-2)/2), (kh-2) mod 2, floor((kh-2)/2)); zero padding is the TMA
//            out-of-bounds fill.
//   deconv : 4 sub-pixel phases (py, px); each is a stride-1 conv over the INPUT grid with 3x3 /
//            3x2 / 2x3 / 2x2 taps (oh = 2 ih - 2 + kh  =>  kh = py (mod 2), ih = m + (py + 2 - kh)/2),
//            its output interleaved at (2m + py, 2n + px).
//   A tile = 128 pixels (bw x bh x nb box) x one swizzle row of K (128/64/32 bytes), landed by one
//            TMA into the canonical K-major swizzled layout the UMMA descriptor expects.
//   B tile = BLOCK_N x the same K slice of the BatchNorm-folded, pre-packed weights.
//   Epilogue: TMEM -> registers (tcgen05.ld) -> bias + LeakyReLU/ReLU -> bf16/fp32 -> stored at
//            the channel offset of the consumer's concat buffer (torch.cat as an addressing mode).
//   Split-K (deep layers, few pixels, many taps): partial fp32 tiles + a deterministic reduction.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#include "unet_internal.cuh"
#include "tc_ptx.cuh"
#include "tc_conv_common.cuh"

#include <cudaTypedefs.h>
#include <type_traits>
#include <cstdlib>
#include <mutex>

namespace svs {


template <int kBlockN, int kSwz, int kStages>
constexpr size_t tc_smem_bytes() {
  return static_cast<size_t>(kStages) * (128 + kBlockN) * kSwz + 1024 /*alignment slack*/ + 256 /*barriers*/ +
         2048 /*bias*/;
}



// Persistent kernel: each CTA walks tiles  blockIdx.x, blockIdx.x + gridDim.x, ...  The accumulator is
// double buffered in TMEM so the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1.
// tile index -> (z = phase * split_k + split, n tile, m tile) with m fastest (neighbouring CTAs share
// the same weight tile in L2).
template <typename OutT, bool kTf32, int kBlockN, int kSwz, int kStages>
__global__ void __launch_bounds__(kTcThreads)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b0,
               const __grid_constant__ CUtensorMap tmap_b1, const __grid_constant__ CUtensorMap tmap_b2,
               const __grid_constant__ CUtensorMap tmap_b3, const TcParams p) {
  constexpr int kABytes = 128 * kSwz;
  constexpr int kBBytes = kBlockN * kSwz;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kAccCols = kBlockN < 32 ? 32 : kBlockN;  // columns per accumulator stage
  constexpr int kTmemCols = 2 * kAccCols;
  constexpr int kMmaPerChunk = kSwz / 32;          // UMMA_K is 32 bytes for bf16 (16) and tf32 (8)

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
  float* sbias = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes + 256);   // [bias_n] staged once
  const int bias_n = p.merged ? p.cout_phase : p.cout;
  for (int i = threadIdx.x; i < bias_n; i += kTcThreads) sbias[i] = __ldg(&p.bias[i]);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = p.m_tiles, n_tiles = p.n_tiles;
  const int per_z = m_tiles * n_tiles;
  const int total_tiles = per_z * p.n_phases * p.split_k;

  long long* dbg = p.dbg ? p.dbg + 8 * blockIdx.x : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = dbg_now();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 4); }
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
  pdl_wait();                                     // previous layer complete: activations readable, outputs writable

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the loop (uniform control flow), one elected lane issues =====
    {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int z = tile / per_z, rem = tile - z * per_z;
        const int nt = rem / m_tiles, mt = rem - nt * m_tiles;
        const int phase = z / p.split_k, split = z - phase * p.split_k;
        const int n_chunks = p.n_chunks[phase];
        const int per_split = (n_chunks + p.split_k - 1) / p.split_k;
        const int c_begin = split * per_split;
        const int c_end = min(n_chunks, c_begin + per_split);
        const int tw = mt % p.ntw, th = (mt / p.ntw) % p.nth, tb = mt / (p.ntw * p.nth);
        const CUtensorMap* tb_map = phase == 0 ? &tmap_b0 : phase == 1 ? &tmap_b1 : phase == 2 ? &tmap_b2 : &tmap_b3;
        const TcChunk* chunks = p.chunks + p.chunk_begin[phase];
        for (int ci = c_begin; ci < c_end; ++ci, ++it) {
          const int s = it % kStages;
          const uint32_t par = (it / kStages) & 1;
          mbar_wait(empty_bar(s), par ^ 1);
          const TcChunk ch = chunks[ci];
          const uint32_t a_dst = smem_base + s * kStageBytes;
          if (elect_one_sync()) {
            mbar_expect_tx(full_bar(s), kStageBytes);
            tma_load_5d(a_dst, &tmap_a, full_bar(s), ch.c_inner, tw * p.bw + ch.dw, ch.ph, th * p.bh + ch.dh,
                        tb * p.nb);
            tma_load_2d(a_dst + kABytes, tb_map, full_bar(s), ci * p.block_k, nt * kBlockN);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();                               // reconverge before the aligned block barrier below
  } else if (warp == 1) {
    // ===== MMA issuer: uniform control flow, one elected lane issues tcgen05.mma / commit =====
    {
      constexpr uint32_t idesc = make_idesc<kTf32, kBlockN>();
      int it = 0, t = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
        const int z = tile / per_z;
        const int phase = z / p.split_k, split = z - phase * p.split_k;
        const int n_chunks = p.n_chunks[phase];
        const int per_split = (n_chunks + p.split_k - 1) / p.split_k;
        const int c_begin = split * per_split;
        const int n_iter = max(0, min(n_chunks, c_begin + per_split) - c_begin);
        const int as = t & 1;
        mbar_wait(tmem_empty_bar(as), ((t >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kAccCols;
        for (int i = 0; i < n_iter; ++i, ++it) {
          const int s = it % kStages;
          const uint32_t par = (it / kStages) & 1;
          mbar_wait(full_bar(s), par);
          if (dbg && it == 0 && lane == 0) dbg[2] = dbg_now();
          tc_fence_after();
          // dispatch on the stage index so that descriptors are "uniform base + compile-time constant":
          // descriptor arithmetic in the vector datapath costs an R2UR chain per MMA (see zc_conv.cu)
          const uint32_t acc0 = i > 0 ? 1u : 0u;
          dispatch_stage<0, kStages>(s, [&](auto sc) {
            constexpr int S = decltype(sc)::value;
            const uint32_t a_addr = smem_base + S * kStageBytes;
            const uint64_t da = make_smem_desc<kSwz>(a_addr);
            const uint64_t db = make_smem_desc<kSwz>(a_addr + kABytes);
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < kMmaPerChunk; ++k) {
                // advance 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
                umma<kTf32>(tmem_d, da + 2u * k, db + 2u * k, idesc, k > 0 ? 1u : acc0);
              }
              umma_commit(bar_base + 8u * (kStages + S));   // empty_bar(S): frees the stage when the MMAs retire
            }
            __syncwarp();
          });
        }
        if (elect_one_sync()) umma_commit(tmem_full_bar(as));   // accumulator complete
        __syncwarp();
        if (dbg && t == 0 && lane == 0) dbg[3] = dbg_now();
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int r = 32 * q + lane;
    const int iw = r % p.bw;
    const int ih = (r / p.bw) % p.bh;
    const int ib = r / (p.bw * p.bh);
    int t = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
      const int z = tile / per_z, rem = tile - z * per_z;
      const int nt = rem / m_tiles, mt = rem - nt * m_tiles;
      const int phase = z / p.split_k, split = z - phase * p.split_k;
      const int n_chunks = p.n_chunks[phase];
      const int per_split = (n_chunks + p.split_k - 1) / p.split_k;
      const int c_begin = split * per_split;
      const int n_iter = max(0, min(n_chunks, c_begin + per_split) - c_begin);
      const int tw = mt % p.ntw, th = (mt / p.ntw) % p.nth, tb = mt / (p.ntw * p.nth);
      const int n0 = nt * kBlockN;
      const int gx = tw * p.bw + iw, gy = th * p.bh + ih, b = tb * p.nb + ib;
      const bool valid = b < p.batch;
      const int as = t & 1;
      mbar_wait(tmem_full_bar(as), (t >> 1) & 1);
      if (dbg && t == 0 && threadIdx.x == 64) dbg[4] = dbg_now();
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * kAccCols;
      if (p.split_k == 1) {
        OutT* const out_base = reinterpret_cast<OutT*>(p.out);
        constexpr int kStep = kBlockN >= 32 ? 32 : 16;
#pragma unroll 2
        for (int c = 0; c < kBlockN; c += kStep) {
          uint32_t v[kStep];
          if (n_iter > 0) {
            tmem_ld16(taddr + c, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
            if constexpr (kStep == 32) tmem_ld16(taddr + c + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < kStep; ++i) v[i] = 0u;
          }
#pragma unroll
          for (int h = 0; h < kStep; h += 16) {
            // merged deconv: 16-column groups never straddle a phase (channels per phase >= 16)
            int py = p.py[phase], px = p.px[phase], ch = n0 + c + h;
            if (p.merged) {
              const int ph = ch / p.cout_phase;
              ch -= ph * p.cout_phase;
              py = ph >> 1; px = ph & 1;
            }
            const int oy = gy * p.out_scale + py, ox = gx * p.out_scale + px;
            OutT* dst = out_base + ((static_cast<size_t>(b) * p.hout + oy) * p.wout + ox) * p.out_pitch + p.out_coff + ch;
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(&sbias[ch + i]);
              f[i] = tc_act(__uint_as_float(v[h + i]) + bv.x, p.act);
              f[i + 1] = tc_act(__uint_as_float(v[h + i + 1]) + bv.y, p.act);
              f[i + 2] = tc_act(__uint_as_float(v[h + i + 2]) + bv.z, p.act);
              f[i + 3] = tc_act(__uint_as_float(v[h + i + 3]) + bv.w, p.act);
            }
            if constexpr (kTf32) {
              if ((p.out_flags & OUT_ACCUMULATE) && valid) {          // data gradients meeting at a skip connection
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 o = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dst) + i);
                  f[i] += o.x; f[i + 1] += o.y; f[i + 2] += o.z; f[i + 3] += o.w;
                }
              }
              if (!(p.out_flags & OUT_KEEP_FP32)) {
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = round_tf32(f[i]);
              }
            }
            if (valid) store16(dst, f);
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
      if (lane == 0) mbar_arrive(tmem_empty_bar(as));      // 4 warps -> accumulator stage is free
      if (dbg && t == 0 && threadIdx.x == 64) dbg[5] = dbg_now();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[6] = dbg_now();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// Deterministic split-K reduction + bias + activation + channel-offset store.
template <typename OutT>
__global__ void __launch_bounds__(256)
splitk_finish_kernel(const TcParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int cg_n = p.cout >> 2;
  const size_t total = static_cast<size_t>(p.n_phases) * p.m_pad * cg_n;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % cg_n);
  const int m = static_cast<int>((idx / cg_n) % p.m_pad);
  const int phase = static_cast<int>(idx / (static_cast<size_t>(cg_n) * p.m_pad));
  const int tile = m >> 7, r = m & 127;
  const int tw = tile % p.ntw, th = (tile / p.ntw) % p.nth, tb = tile / (p.ntw * p.nth);
  const int iw = r % p.bw, ih = (r / p.bw) % p.bh, ib = r / (p.bw * p.bh);
  const int gx = tw * p.bw + iw, gy = th * p.bh + ih, b = tb * p.nb + ib;
  if (b >= p.batch) return;
  const int n = cg * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < p.split_k; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(
        p.partial + (static_cast<size_t>(phase * p.split_k + s) * p.m_pad + m) * p.cout + n);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float o[4] = {tc_act(acc.x + p.bias[n], p.act), tc_act(acc.y + p.bias[n + 1], p.act),
                      tc_act(acc.z + p.bias[n + 2], p.act), tc_act(acc.w + p.bias[n + 3], p.act)};
  const int oy = gy * p.out_scale + p.py[phase], ox = gx * p.out_scale + p.px[phase];
  OutT* dst = reinterpret_cast<OutT*>(p.out) +
              ((static_cast<size_t>(b) * p.hout + oy) * p.wout + ox) * p.out_pitch + p.out_coff + n;
  if constexpr (sizeof(OutT) == 2) {
    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
    __nv_bfloat162 c = __floats2bfloat162_rn(o[2], o[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&c);
    *reinterpret_cast<uint2*>(dst) = u;
  } else {
    float4 r = make_float4(o[0], o[1], o[2], o[3]);
    if (p.out_flags & OUT_ACCUMULATE) {
      const float4 old = *reinterpret_cast<const float4*>(dst);
      r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
    }
    if (!(p.out_flags & OUT_KEEP_FP32)) r = make_float4(round_tf32(r.x), round_tf32(r.y), round_tf32(r.z), round_tf32(r.w));
    *reinterpret_cast<float4*>(dst) = r;
  }
}

// [tap][ci][co] fp32 (BN folded) -> [co][K] K-major in chunk order, bf16 or fp32
template <typename E>
__global__ void tc_pack_weights_kernel(const float* __restrict__ w_fold, int cin, int cout,
                                       const int2* __restrict__ chunk_src, int n_chunks, int bk,
                                       E* __restrict__ out) {
  const int k_total = n_chunks * bk;
  const size_t total = static_cast<size_t>(cout) * k_total;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / k_total);
    const int k = static_cast<int>(i % k_total);
    const int2 src = chunk_src[k / bk];
    const float v = w_fold[(static_cast<size_t>(src.x) * cin + src.y + (k % bk)) * cout + n];
    if constexpr (sizeof(E) == 2) out[i] = __float2bfloat16_rn(v);
    else out[i] = round_tf32(v);
  }
}

// merged deconv: B[n = phase*cout + co][k = chunk*bk + kk]; chunk_src = ((dh+1)*3 + (dw+1), ci0);
// phase (py, px) uses tap (kh, kw) = (py + 2 - 2 dh, px + 2 - 2 dw) when it lies in [0, 4], else zero.
template <typename E>
__global__ void tc_pack_weights_merged_kernel(const float* __restrict__ w_fold, int cin, int cout,
                                              const int2* __restrict__ chunk_src, int n_chunks, int bk,
                                              E* __restrict__ out) {
  const int k_total = n_chunks * bk;
  const size_t total = static_cast<size_t>(4 * cout) * k_total;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / k_total);
    const int k = static_cast<int>(i % k_total);
    const int ph = n / cout, co = n % cout;
    const int py = ph >> 1, px = ph & 1;
    const int2 src = chunk_src[k / bk];
    const int dh = src.x / 3 - 1, dw = src.x % 3 - 1;
    const int kh = py + 2 - 2 * dh, kw = px + 2 - 2 * dw;
    float v = 0.0f;
    if (kh >= 0 && kh <= 4 && kw >= 0 && kw <= 4)
      v = w_fold[(static_cast<size_t>(kh * 5 + kw) * cin + src.y + (k % bk)) * cout + co];
    if constexpr (sizeof(E) == 2) out[i] = __float2bfloat16_rn(v);
    else out[i] = round_tf32(v);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
static PFN_cuTensorMapEncodeTiled get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  });
  return fn;
}

int encode_tensor_map(CUtensorMap* map, bool tf32, int rank, void* base, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box, int swz, int l2_promotion) {
  PFN_cuTensorMapEncodeTiled fn = get_encode_fn();
  if (!fn) return fail(SVS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  // swz 12832 = 128-byte span swizzled in 32-byte atoms (the only layout tcgen05 accepts for MN-major TF32 operands)
  const CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                : swz == 12832 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base,
                  dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  l2_promotion >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                  : l2_promotion >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                  : l2_promotion >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SVS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  return SVS_OK;
}

static int make_tmap_a(const TcLayer& t, const void* buf, int batch, int es, bool tf32, CUtensorMap* out) {
  const ConvDesc& g = t.d;
  const cuuint64_t ct = g.in_pitch;
  const cuuint64_t H = g.hin, W = g.win;
  cuuint64_t dims[5], strides[4];
  if (!g.transposed) {
    dims[0] = 2 * ct; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = batch;
    strides[0] = 2 * ct * es; strides[1] = W * ct * es; strides[2] = 2 * W * ct * es; strides[3] = H * W * ct * es;
  } else {
    dims[0] = ct; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = batch;
    strides[0] = ct * es; strides[1] = W * ct * es; strides[2] = W * ct * es; strides[3] = H * W * ct * es;
  }
  const cuuint32_t box[5] = {static_cast<cuuint32_t>(t.block_k), static_cast<cuuint32_t>(t.bw), 1,
                             static_cast<cuuint32_t>(t.bh), static_cast<cuuint32_t>(t.nb)};
  return encode_tensor_map(out, tf32, 5, const_cast<void*>(buf), dims, strides, box, t.swz);
}

static unsigned tc_disable_mask() {
  const char* e = std::getenv("SVS_TC_DISABLE_MASK");   // debugging: bit li set -> layer li on the CUDA-core path
  return e ? static_cast<unsigned>(std::strtoul(e, nullptr, 0)) : 0u;
}

// (Re)packs the weights of a planned problem from w_fold [25][cin][cout] fp32 (stream ordered; the training step
// calls this every iteration, the inference plan once).
int tc_pack_one(TcLayer& t, const float* w_fold, bool tf32, cudaStream_t st) {
  const ConvDesc& g = t.d;
  const int n_total = t.merged ? 4 * g.cout : g.cout;
  for (int ph = 0; ph < t.n_phases; ++ph) {
    const TcPhase& phs = t.phases[ph];
    const size_t n = static_cast<size_t>(n_total) * t.k_total[ph];
    const size_t off = static_cast<size_t>(phs.b_elem_off);
    const unsigned blocks = static_cast<unsigned>((n + 255) / 256 > 2368 ? 2368 : (n + 255) / 256);
    if (t.merged && tf32)
      tc_pack_weights_merged_kernel<float><<<blocks, 256, 0, st>>>(w_fold, g.cin, g.cout, t.d_src, phs.n_chunks,
                                                                  t.block_k, static_cast<float*>(t.d_weights));
    else if (t.merged)
      tc_pack_weights_merged_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
          w_fold, g.cin, g.cout, t.d_src, phs.n_chunks, t.block_k, static_cast<__nv_bfloat16*>(t.d_weights));
    else if (tf32)
      tc_pack_weights_kernel<float><<<blocks, 256, 0, st>>>(w_fold, g.cin, g.cout, t.d_src + phs.chunk_begin,
                                                           phs.n_chunks, t.block_k,
                                                           static_cast<float*>(t.d_weights) + off);
    else
      tc_pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
          w_fold, g.cin, g.cout, t.d_src + phs.chunk_begin, phs.n_chunks, t.block_k,
          static_cast<__nv_bfloat16*>(t.d_weights) + off);
    SVS_CHECK_LAUNCH("tc_pack_weights_kernel");
  }
  return SVS_OK;
}

// Plans one problem: tiling, K-chunk schedule, device tables, weight buffer + tensor maps.  `w_fold` may be null
// (weights are packed later with tc_pack_one).  Leaves t.enabled == false when the shape has no tcgen05 mapping.
int tc_plan_one(TcLayer& t, const ConvDesc& g, const float* w_fold, bool tf32, cudaStream_t st) {
  const int es = tf32 ? 4 : 2;
  t.d = g;
  t.enabled = false;
  int swz = g.cin * es >= 128 ? 128 : g.cin * es;     // one swizzle row = min(128 B, all input channels)
  if (swz != 32 && swz != 64 && swz != 128) return SVS_OK;
  t.swz = swz;
  t.block_k = swz / es;
  const char* no_merge = std::getenv("SVS_TC_NO_MERGE");
  t.merged = g.transposed && 4 * g.cout <= 256 && !(no_merge && no_merge[0] == '1');
  const int n_total = t.merged ? 4 * g.cout : g.cout;
  t.block_n = n_total < 256 ? (n_total < 128 ? n_total : 128) : (t.merged ? 256 : 128);
  if (n_total % t.block_n != 0 || (t.block_n != 16 && t.block_n != 32 && t.block_n != 64 && t.block_n != 128 &&
                                   t.block_n != 256))
    return SVS_OK;
  t.gw = g.transposed ? g.win : g.wout;
  t.gh = g.transposed ? g.hin : g.hout;
  t.bw = t.gw < 16 ? t.gw : 16;
  t.bh = t.gh < 128 / t.bw ? t.gh : 128 / t.bw;
  t.nb = 128 / (t.bw * t.bh);
  // ---- K-chunk schedule ----
  std::vector<int2> src;
  const int ct = g.in_pitch;
  t.chunks.clear();
  if (!g.transposed) {
    t.n_phases = 1;
    t.phases[0] = TcPhase{0, 0, 0, 0, 0};
    for (int kh = 0; kh < 5; ++kh)
      for (int kw = 0; kw < 5; ++kw) {
        const int qh = kh - 2, qw = kw - 2;
        const int ph = qh & 1, pw = qw & 1;
        const int dh = (qh - ph) / 2, dw = (qw - pw) / 2;
        for (int c0 = 0; c0 < g.cin; c0 += t.block_k) {
          t.chunks.push_back(TcChunk{pw * ct + g.in_coff + c0, dw, ph, dh});
          src.push_back(make_int2(kh * 5 + kw, c0));
        }
      }
    t.phases[0].n_chunks = static_cast<int>(t.chunks.size());
  } else if (t.merged) {
    t.n_phases = 1;
    t.phases[0] = TcPhase{0, 0, 0, 0, 0};
    for (int dh = 1; dh >= -1; --dh)
      for (int dw = 1; dw >= -1; --dw)
        for (int c0 = 0; c0 < g.cin; c0 += t.block_k) {
          t.chunks.push_back(TcChunk{g.in_coff + c0, dw, 0, dh});
          src.push_back(make_int2((dh + 1) * 3 + (dw + 1), c0));
        }
    t.phases[0].n_chunks = static_cast<int>(t.chunks.size());
  } else {
    t.n_phases = 4;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        TcPhase& phs = t.phases[py * 2 + px];
        phs.py = py; phs.px = px;
        phs.chunk_begin = static_cast<int>(t.chunks.size());
        for (int kh = py; kh < 5; kh += 2)
          for (int kw = px; kw < 5; kw += 2) {
            const int dh = (py + 2 - kh) / 2, dw = (px + 2 - kw) / 2;
            for (int c0 = 0; c0 < g.cin; c0 += t.block_k) {
              t.chunks.push_back(TcChunk{g.in_coff + c0, dw, 0, dh});
              src.push_back(make_int2(kh * 5 + kw, c0));
            }
          }
        phs.n_chunks = static_cast<int>(t.chunks.size()) - phs.chunk_begin;
      }
  }
  // ---- upload schedule, weight buffer, B tensor maps ----
  const size_t n_chunks_total = t.chunks.size();
  SVS_CUDA_TRY(cudaMalloc(&t.d_chunks, sizeof(TcChunk) * n_chunks_total));
  SVS_CUDA_TRY(cudaMalloc(&t.d_src, sizeof(int2) * n_chunks_total));
  SVS_CUDA_TRY(cudaMemcpyAsync(t.d_chunks, t.chunks.data(), sizeof(TcChunk) * n_chunks_total,
                               cudaMemcpyHostToDevice, st));
  SVS_CUDA_TRY(cudaMemcpyAsync(t.d_src, src.data(), sizeof(int2) * n_chunks_total, cudaMemcpyHostToDevice, st));
  const size_t w_elems = n_chunks_total * t.block_k * n_total;
  SVS_CUDA_TRY(cudaMalloc(&t.d_weights, w_elems * es));
  size_t off = 0;
  for (int ph = 0; ph < t.n_phases; ++ph) {
    TcPhase& phs = t.phases[ph];
    phs.b_elem_off = static_cast<int64_t>(off);
    t.k_total[ph] = phs.n_chunks * t.block_k;
    const size_t n = static_cast<size_t>(n_total) * t.k_total[ph];
    // B tensor map: [Cout][K] K-major
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(t.k_total[ph]), static_cast<cuuint64_t>(n_total)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(t.k_total[ph]) * es};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(t.block_k), static_cast<cuuint32_t>(t.block_n)};
    int rc = encode_tensor_map(&t.tmap_b[ph], tf32, 2, static_cast<char*>(t.d_weights) + off * es, dims, strides, box,
                               t.swz);
    if (rc != SVS_OK) return rc;
    t.tmap_b_wide[ph] = t.tmap_b[ph];
    if (!t.merged && t.block_n == 128 && t.swz == 128 && n_total % 256 == 0) {
      const cuuint32_t box_wide[2] = {static_cast<cuuint32_t>(t.block_k), 256u};
      rc = encode_tensor_map(&t.tmap_b_wide[ph], tf32, 2, static_cast<char*>(t.d_weights) + off * es, dims, strides,
                             box_wide, t.swz);
      if (rc != SVS_OK) return rc;
      t.has_wide = true;
    }
    off += n;
  }
  for (int ph = t.n_phases; ph < 4; ++ph) { t.tmap_b[ph] = t.tmap_b[0]; t.tmap_b_wide[ph] = t.tmap_b_wide[0]; }
  SVS_CUDA_TRY(cudaStreamSynchronize(st));     // the host vectors are consumed
  if (w_fold) {
    int rc = tc_pack_one(t, w_fold, tf32, st);
    if (rc != SVS_OK) return rc;
  }
  t.enabled = true;
  return SVS_OK;
}

void tc_free_one(TcLayer& t) {
  if (t.d_chunks) cudaFree(t.d_chunks);
  if (t.d_src) cudaFree(t.d_src);
  if (t.d_weights) cudaFree(t.d_weights);
  t.d_chunks = nullptr; t.d_src = nullptr; t.d_weights = nullptr; t.enabled = false;
}

int tc_plan_layers(svs_unet_plan* plan, cudaStream_t st) {
  const bool tf32 = plan->precision == SVS_PRECISION_TF32;
  const unsigned disable = tc_disable_mask();
  for (int li = 1; li <= 10; ++li) {
    TcLayer& t = plan->tc[li];
    t.layer = li;
    if ((disable >> li) & 1u) continue;
    int rc = tc_plan_one(t, desc_of_layer(li), plan->w_fold[li], tf32, st);
    if (rc != SVS_OK) return rc;
  }
  return SVS_OK;
}

void tc_free_layers(svs_unet_plan* plan) {
  for (int li = 0; li < 12; ++li) tc_free_one(plan->tc[li]);
}

// SVS_TC_CLUSTER: 0 = one CTA per tile stream + split-K finish kernel,
// 2 (default) = split-K across a cluster, reduced through distributed shared memory (conv_tc_cluster.cu)
static int ilog2_exact(int v) {      // log2 of a power of two, else -1
  for (int k = 0; k < 31; ++k) if ((1 << k) == v) return k;
  return -1;
}
int tc_cluster_mode() {
  static const int mode = [] { const char* e = std::getenv("SVS_TC_CLUSTER"); return e ? std::atoi(e) : 2; }();
  return mode;
}
bool ck_supported(const TcLayer& t, int split, int block_n);
int ck_launch_layer(const TcLayer& t, bool tf32, int block_n, const CUtensorMap& ta, const TcParams& p, cudaStream_t st);

// Tiling of one problem at one batch size: M tiles, the N tile (a layer with has_wide may use 256) and split-K.
void tc_tiling(const TcLayer& t, int batch, int* m_tiles, int* split_k, int* block_n) {
  const ConvDesc& g = t.d;
  const int ntw = t.gw / t.bw, nth = t.gh / t.bh, ntb = (batch + t.nb - 1) / t.nb;
  *m_tiles = ntw * nth * ntb;
  *block_n = t.block_n;
  const int n_total = t.merged ? 4 * g.cout : g.cout;
  const int tiles = *m_tiles * (n_total / t.block_n) * t.n_phases;
  int min_chunks = 1 << 30;
  for (int ph = 0; ph < t.n_phases; ++ph) min_chunks = t.phases[ph].n_chunks < min_chunks ? t.phases[ph].n_chunks : min_chunks;
  int s = 1;
  const char* e = std::getenv("SVS_TC_SPLITK");
  if (e) s = std::atoi(e);
  else if (tiles < 120) s = 148 / tiles;
  if (s > 8) s = 8;
  if (s > min_chunks / 4) s = min_chunks / 4;
  if (tc_cluster_mode() == 2 && !e) {               // cluster split-K reduces 128 / s rows per CTA
    while (s & (s - 1)) --s;
  }
  if (s < 1 || t.merged) s = 1;
  *split_k = s;
  // These layers are bound by L2 -> shared-memory bytes (~55 B/clk/SM): a 128 x 256 tile moves 48 KB per K chunk
  // for twice the MACs of a 128 x 128 tile's 32 KB.  Worth it once the wider tiles still fill the SMs.
  static const bool no_wide = [] { const char* w = std::getenv("SVS_TC_NO_WIDE"); return w && w[0] == '1'; }();
  if (t.has_wide && s == 1 && !no_wide) {
    const int sms = num_sms();
    const int wide_tiles = *m_tiles * (n_total / 256) * t.n_phases;
    const int cost_narrow = (tiles + sms - 1) / sms * 32, cost_wide = (wide_tiles + sms - 1) / sms * 48;
    if (cost_wide < cost_narrow) *block_n = 256;
  }
  // Cluster split-K on 128 x 256 tiles: half the tiles, twice the split, a quarter less operand traffic per MAC.
  // OFF by default (SVS_CK_WIDE=1): parity-tested, measured at batch 64 bf16 conv5 12.7 vs 12.9 us, deconv1 15.0 vs
  // 15.5 us, conv6 24.9 vs 13.3 us (sixteen clusters of EIGHT 200 KB CTAs do not all find a GPC in one wave).  The
  // halved A-chunk count and the quarter less traffic do not shorten the main loop, i.e. these layers are not bound by
  // bytes into shared memory either; the doubled split makes the DSMEM reduction longer.
  static const bool ck_wide = [] { const char* w = std::getenv("SVS_CK_WIDE"); return w && w[0] == '1'; }();
  if (t.has_wide && s > 1 && tc_cluster_mode() == 2 && !e && !no_wide && ck_wide) {
    const int wide_tiles = *m_tiles * (n_total / 256) * t.n_phases;
    int sw = 148 / wide_tiles;
    if (sw > 8) sw = 8;
    if (sw > min_chunks / 4) sw = min_chunks / 4;
    while (sw & (sw - 1)) --sw;
    if (sw >= 2 && wide_tiles * sw >= tiles * s && ck_supported(t, sw, 256)) {
      *block_n = 256;
      *split_k = sw;
    }
  }
}

size_t tc_splitk_bytes_one(const TcLayer& t, int batch) {
  if (!t.enabled) return 0;
  int m_tiles, split, block_n;
  tc_tiling(t, batch, &m_tiles, &split, &block_n);
  if (split <= 1) return 0;
  return static_cast<size_t>(split) * t.n_phases * m_tiles * 128 * t.d.cout * sizeof(float);
}

size_t tc_splitk_bytes(const svs_unet_plan* plan, int batch) {
  size_t best = 0;
  for (int li = 0; li < 12; ++li) {
    const size_t bytes = tc_splitk_bytes_one(plan->tc[li], batch);
    best = bytes > best ? bytes : best;
  }
  return best;
}

int tc_launch_count(const svs_unet_plan* plan, int li, int batch) {
  int m_tiles, split, block_n;
  tc_tiling(plan->tc[li], batch, &m_tiles, &split, &block_n);
  if (split > 1 && tc_cluster_mode() == 2 && ck_supported(plan->tc[li], split, block_n)) return 1;
  return split > 1 ? 2 : 1;   // main kernel (+ split-K reduction)
}

template <typename OutT, bool kTf32, int kBlockN, int kSwz, int kStages>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap* tb, const TcParams& p, int total_tiles, cudaStream_t st) {
  auto kern = tc_conv_kernel<OutT, kTf32, kBlockN, kSwz, kStages>;
  constexpr size_t smem = tc_smem_bytes<kBlockN, kSwz, kStages>();
  SVS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  // resident CTAs per SM: shared memory, TMEM columns (2 accumulator stages) and a cap of 3
  constexpr int kAccCols = kBlockN < 32 ? 32 : kBlockN;
  int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  if (per_sm > 512 / (2 * kAccCols)) per_sm = 512 / (2 * kAccCols);
  if (per_sm > 3) per_sm = 3;
  if (per_sm < 1) per_sm = 1;
  int grid = num_sms() * per_sm;
  if (grid > total_tiles) grid = total_tiles;
  SVS_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(kTcThreads), smem, st, ta, tb[0], tb[1], tb[2], tb[3], p));
  return SVS_OK;
}

long long* g_tc_dbg = nullptr;   // set through svs_debug_set_trace (profiling only)
int g_tc_dbg_layer = -1;

// Enqueues one planned problem on `io`'s buffers.
int tc_launch(const TcLayer& t, const TcIo& io, int batch, bool tf32, cudaStream_t st) {
  const ConvDesc& g = t.d;
  const int es = tf32 ? 4 : 2;
  static std::mutex mu;
  CUtensorMap ta;
  {
    std::lock_guard<std::mutex> lock(mu);
    int hit = -1;
    for (int i = 0; i < TcLayer::kTmapCache; ++i)
      if (t.tmap_a_base[i] == io.in && t.tmap_a_batch[i] == batch) hit = i;
    if (hit < 0) {
      hit = t.tmap_a_next;
      t.tmap_a_next = (t.tmap_a_next + 1) % TcLayer::kTmapCache;
      t.tmap_a_base[hit] = nullptr;
      int rc = make_tmap_a(t, io.in, batch, es, tf32, &t.tmap_a[hit]);
      if (rc != SVS_OK) return rc;
      t.tmap_a_base[hit] = io.in;
      t.tmap_a_batch[hit] = batch;
    }
    ta = t.tmap_a[hit];
  }
  TcParams p{};
  p.chunks = t.d_chunks;
  for (int ph = 0; ph < 4; ++ph) {
    p.n_chunks[ph] = t.phases[ph < t.n_phases ? ph : 0].n_chunks;
    p.chunk_begin[ph] = t.phases[ph < t.n_phases ? ph : 0].chunk_begin;
    p.py[ph] = t.phases[ph < t.n_phases ? ph : 0].py;
    p.px[ph] = t.phases[ph < t.n_phases ? ph : 0].px;
  }
  int m_tiles, split, block_n;
  tc_tiling(t, batch, &m_tiles, &split, &block_n);
  const CUtensorMap* tb = block_n == t.block_n ? t.tmap_b : t.tmap_b_wide;
  p.n_phases = t.n_phases;
  p.split_k = split;
  p.ntw = t.gw / t.bw; p.nth = t.gh / t.bh;
  p.bw = t.bw; p.bh = t.bh; p.nb = t.nb;
  p.bw_log2 = ilog2_exact(t.bw); p.bh_log2 = ilog2_exact(t.bh);
  p.batch = batch;
  p.block_k = t.block_k;
  p.out = io.out;
  p.out_pitch = g.out_pitch;
  p.out_coff = g.out_coff;
  p.hout = g.hout; p.wout = g.wout;
  p.out_scale = g.transposed ? 2 : 1;
  p.bias = io.bias;
  p.act = g.act;
  p.out_flags = io.out_flags;
  p.partial = io.splitk;
  p.m_pad = m_tiles * 128;
  p.cout = t.merged ? 4 * g.cout : g.cout;
  p.merged = t.merged ? 1 : 0;
  p.cout_phase = g.cout;
  p.dbg = io.dbg;
  p.m_tiles = m_tiles;
  p.n_tiles = p.cout / block_n;
  const int grid = m_tiles * p.n_tiles * t.n_phases * split;
  int rc = SVS_ERR_NOT_IMPLEMENTED;
  bool finish = split > 1;
  if (tc_cluster_mode() == 2 && split > 1 && ck_supported(t, split, block_n)) {
    rc = ck_launch_layer(t, tf32, block_n, ta, p, st);  // split-K inside a cluster: no partials, no finish kernel
    finish = false;
  } else if (split > 1 && io.splitk_bytes < static_cast<size_t>(split) * t.n_phases * p.m_pad * g.cout * sizeof(float)) {
    return fail(SVS_ERR_WORKSPACE, "tc_launch: split-K scratch too small");
  }
#define SVS_TC_CASE(N, S, ST)                                                                       \
  if (rc == SVS_ERR_NOT_IMPLEMENTED && block_n == N && t.swz == S) { \
    rc = tf32 ? launch_tc<float, true, N, S, ST>(ta, tb, p, grid, st)                                \
              : launch_tc<__nv_bfloat16, false, N, S, ST>(ta, tb, p, grid, st);                      \
  }
  SVS_TC_CASE(256, 128, 4)
  SVS_TC_CASE(128, 128, 6)
  SVS_TC_CASE(64, 128, 4)
  SVS_TC_CASE(32, 128, 4)
  SVS_TC_CASE(16, 128, 4)
  SVS_TC_CASE(128, 64, 4)
  SVS_TC_CASE(64, 64, 4)
  SVS_TC_CASE(32, 64, 4)
  SVS_TC_CASE(32, 32, 4)
#undef SVS_TC_CASE
  if (rc == SVS_ERR_NOT_IMPLEMENTED)
    return fail(rc, "tc_launch: no kernel instantiation for block_n=" + std::to_string(block_n) +
                        " swz=" + std::to_string(t.swz));
  if (rc != SVS_OK) return rc;
  if (finish) {
    const size_t total = static_cast<size_t>(t.n_phases) * p.m_pad * (g.cout / 4);
    const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
    if (tf32) SVS_CUDA_TRY(launch_pdl(splitk_finish_kernel<float>, dim3(blocks), dim3(256), 0, st, p));
    else SVS_CUDA_TRY(launch_pdl(splitk_finish_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, p));
  }
  return SVS_OK;
}

int tc_launch_layer(const svs_unet_plan* plan, int li, const Workspace& ws, int batch, cudaStream_t st) {
  const LayerGeom& g = kLayers[li];
  TcIo io;
  io.in = ws.buf[g.in_buf];
  io.out = ws.buf[g.out_buf];
  io.bias = plan->b_fold[li];
  io.splitk = ws.splitk;
  io.splitk_bytes = ws.splitk_bytes;
  io.dbg = (g_tc_dbg_layer == li) ? g_tc_dbg : nullptr;
  return tc_launch(plan->tc[li], io, batch, plan->precision == SVS_PRECISION_TF32, st);
}

}  // namespace svs

extern "C" int svs_debug_set_trace(long long* device_buffer, int layer) {
  svs::g_tc_dbg = device_buffer;
  svs::g_tc_dbg_layer = device_buffer ? layer : -1;
  return SVS_OK;
}
