// D6 + sigmoid + mask x mixture on the tensor cores (reference model.py:109,198,200 and
// inference.py:102,107).
//
// deconv6 has ONE output channel, so the implicit GEMM over output pixels would have N = 1.  It is
// turned inside out instead:
//     P[input pixel][tap] = sum_ci  x[pixel][ci] * w[ci][tap]          (GEMM: M = pixels, N = 25 -> 32, K = 32)
//     y[2m+py, 2n+px]     = sum_{kh = py (2), kw = px (2)}  P[(m + (py+2-kh)/2, n + (px+2-kw)/2)][kh*5+kw]
// i.e. one tcgen05 GEMM per input slab followed by a col2im gather out of shared memory, then
// bias + sigmoid (+ 1-m) (x mixture) and the store into the caller's patch view.
//
// One CTA = one patch x 8 consecutive input rows (1 halo row each side, 6 interior rows = 12 output
// rows x 128 frames).  TMA lands the [8 rows x 64 cols] x 32-channel slab as 512 K-major rows (4 MMA
// M-tiles); the 4 accumulators (4 x 32 TMEM columns) are drained to P[25][512] in shared memory,
// which ALIASES the operand slab (dead once the MMAs have committed).
#include "unet_internal.cuh"
#include "tc_ptx.cuh"

namespace svs {

constexpr int kD6Threads = 256;          // warp 0: TMA + MMA issue, warps 4..7: TMEM drain, all: gather
constexpr int kD6Rows = 8;               // input rows per CTA (incl. halo)
constexpr int kD6Interior = kD6Rows - 2;
constexpr int kD6Pix = kD6Rows * 64;     // 512 GEMM rows
constexpr int kD6PPitch = kD6Pix + 16;   // tap-plane pitch (= 16 mod 32 banks: px=0/px=1 lanes do not collide)
constexpr int kD6Taps = 25;

template <int kSwz>
constexpr size_t d6_smem_bytes() {
  const size_t operands = static_cast<size_t>(kD6Pix) * kSwz + 32 * kSwz;
  const size_t p = sizeof(float) * kD6Taps * kD6PPitch;
  return (operands > p ? operands : p) + 1024 + 64;
}

// col2im for one PAIR of adjacent outputs (2n, 2n+1) of output row 2m+PY: the five kw taps of a kernel row share
// one base address (lane = n: conflict-free), PY fixes kh, so every tap offset is a compile-time constant.
template <int PY>
__device__ __forceinline__ float2 d6_gather_pair(const float* __restrict__ P, int mr, int n, float bs) {
  float a0 = bs, a1 = bs;
  const bool lo = n > 0, hi = n < 63;                                  // columns n-1 / n+1 exist
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int kh = PY + 2 * a;
    if (kh > 4) continue;
    const float* row = P + (mr + ((PY + 2 - kh) >> 1)) * 64 + n;       // halo rows are zero-filled by TMA
    // px = 0: kw = 0, 2, 4 -> columns n+1, n, n-1;  px = 1: kw = 1, 3 -> columns n+1, n
    const float t0 = hi ? row[(kh * 5 + 0) * kD6PPitch + 1] : 0.0f;
    const float t2 = row[(kh * 5 + 2) * kD6PPitch];
    const float t4 = lo ? row[(kh * 5 + 4) * kD6PPitch - 1] : 0.0f;
    const float t1 = hi ? row[(kh * 5 + 1) * kD6PPitch + 1] : 0.0f;
    const float t3 = row[(kh * 5 + 3) * kD6PPitch];
    a0 += t0; a0 += t2; a0 += t4;
    a1 += t1; a1 += t3;
  }
  return make_float2(a0, a1);
}

// kDense: mixture and output are contiguous [B][512][128] float32 (the staged pipeline and the dense UNet entry):
// 32-bit index math, 8-byte mixture loads and stores, a thread owns output pairs.  Otherwise the generic strided
// patch view (one output per thread step, 64-bit strides).
template <bool kTf32, int kSwz, bool kDense>
__global__ void __launch_bounds__(kD6Threads)
deconv6_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const float* __restrict__ bias, const float* __restrict__ mix,
                  const int64_t* __restrict__ mix_off, int64_t mix_sb, int64_t mix_sf, int64_t mix_st,
                  float* __restrict__ out, const int64_t* __restrict__ out_off, int64_t out_sb, int64_t out_sf,
                  int64_t out_st, const int32_t* __restrict__ in_frames, int flags) {
  constexpr int kABytes = kD6Pix * kSwz;
  constexpr int kWBytes = 32 * kSwz;
  constexpr int kKSteps = kSwz / 32;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* P = reinterpret_cast<float*>(smem_gen);                       // aliases the operand slab
  constexpr size_t kOperandEnd = (static_cast<size_t>(kABytes + kWBytes) > sizeof(float) * kD6Taps * kD6PPitch)
                                     ? static_cast<size_t>(kABytes + kWBytes)
                                     : sizeof(float) * kD6Taps * kD6PPitch;
  const uint32_t bar_full = smem_base + static_cast<uint32_t>(kOperandEnd);
  const uint32_t bar_mma = bar_full + 8;
  const uint32_t tmem_slot = bar_full + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kOperandEnd + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * kD6Interior - 1;                       // first input row of the slab (may be -1)

  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(bar_full, kABytes + kWBytes);
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
          " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_base),
          "l"(reinterpret_cast<uint64_t>(&tmap_a)), "r"(bar_full), "r"(0), "r"(0), "r"(row0), "r"(b)
          : "memory");
      tma_load_2d(smem_base + kABytes, &tmap_w, bar_full, 0, 0);
    }
    __syncwarp();
  }
  // ---- prefetch the mixture values this thread will need in the gather (independent of the GEMM), so
  //      their DRAM latency hides behind TMA + MMA + drain instead of serialising inside the gather loop
  constexpr int kOutPerThread = 2 * kD6Interior * 128 / kD6Threads;     // 6
  constexpr int kPairsPerThread = kOutPerThread / 2;                    // 3
  const int nf = in_frames ? in_frames[b] : SVS_PATCH_FRAMES;
  const float* mix_b = mix + (mix_off ? mix_off[b] : b * mix_sb);
  float mixv[kOutPerThread];
  if constexpr (kDense) {
    // pair pr = threadIdx.x + 256 j: n = pr & 63, output row (local) = pr >> 6
#pragma unroll
    for (int j = 0; j < kPairsPerThread; ++j) {
      const int pr = threadIdx.x + j * kD6Threads;
      const int n = pr & 63, oyl = pr >> 6;
      const int oy = 2 * (row0 + 1 + (oyl >> 1)) + (oyl & 1);
      float2 v = make_float2(1.0f, 1.0f);
      // the mixture and the result are touched once per forward: streaming accesses keep them from evicting the
      // skip activations that are still waiting in L2
      if ((flags & SVS_FLAG_APPLY_MASK) && oy < 512) v = __ldcs(reinterpret_cast<const float2*>(mix_b + oy * 128 + 2 * n));
      mixv[2 * j] = v.x; mixv[2 * j + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kOutPerThread; ++j) {
      const int o = threadIdx.x + j * kD6Threads;
      const int ox = o & 127, oyl = o >> 7;
      const int m = row0 + 1 + (oyl >> 1);
      const int oy = 2 * m + (oyl & 1);
      mixv[j] = ((flags & SVS_FLAG_APPLY_MASK) && m < 256 && ox < nf) ? __ldg(mix_b + oy * mix_sf + ox * mix_st) : 1.0f;
    }
  }
  if (warp == 0) {
    mbar_wait(bar_full, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc<kTf32, 32>();
    const uint64_t dw = make_smem_desc<kSwz>(smem_base + kABytes);
    if (elect_one_sync()) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const uint64_t da = make_smem_desc<kSwz>(smem_base + mt * 128 * kSwz);
#pragma unroll
        for (int k = 0; k < kKSteps; ++k) umma<kTf32>(tmem_base + mt * 32, da + 2u * k, dw + 2u * k, idesc, k > 0 ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    __syncwarp();
  }
  if (warp >= 4) {
    // drain the 4 accumulators into P[tap][pixel]; the MMAs have finished reading the operand slab
    const int q = warp & 3;
    mbar_wait(bar_mma, 0);
    tc_fence_after();
#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      const int pix = mt * 128 + 32 * q + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + mt * 32;
      uint32_t v0[16], v1[16];
      tmem_ld16(taddr, v0);
      tmem_ld16(taddr + 16, v1);
      tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 16; ++t) P[t * kD6PPitch + pix] = __uint_as_float(v0[t]);
#pragma unroll
      for (int t = 16; t < kD6Taps; ++t) P[t * kD6PPitch + pix] = __uint_as_float(v1[t - 16]);
    }
    tc_fence_before();
  }
  __syncthreads();

  // ---- col2im gather + sigmoid + mask: 12 output rows x 128 frames ----
  const float bs = __ldg(bias);
  float* out_b = out + (out_off ? out_off[b] : b * out_sb);
  if constexpr (kDense) {
#pragma unroll
    for (int j = 0; j < kPairsPerThread; ++j) {
      const int pr = threadIdx.x + j * kD6Threads;
      const int n = pr & 63, oyl = pr >> 6;                            // oyl parity is warp-uniform
      const int mr = 1 + (oyl >> 1);
      const int oy = 2 * (row0 + mr) + (oyl & 1);
      if (oy >= 512) continue;
      const float2 acc = (oyl & 1) ? d6_gather_pair<1>(P, mr, n, bs) : d6_gather_pair<0>(P, mr, n, bs);
      float m0 = __fdividef(1.0f, 1.0f + __expf(-acc.x));              // torch.sigmoid, model.py:200
      float m1 = __fdividef(1.0f, 1.0f + __expf(-acc.y));
      if (flags & SVS_FLAG_INVERT) { m0 = 1.0f - m0; m1 = 1.0f - m1; } // inference.py:102
      float* dst = out_b + oy * 128 + 2 * n;
      // inference.py:107 (mixv = 1 without APPLY_MASK); frames >= nf are cropped (inference.py:113)
      if (2 * n + 1 < nf) __stcs(reinterpret_cast<float2*>(dst), make_float2(m0 * mixv[2 * j], m1 * mixv[2 * j + 1]));
      else if (2 * n < nf) dst[0] = m0 * mixv[2 * j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < kOutPerThread; ++j) {
      const int o = threadIdx.x + j * kD6Threads;
      const int ox = o & 127, oyl = o >> 7;
      const int mr = 1 + (oyl >> 1);                                     // slab row of the input pixel
      const int m = row0 + mr;                                           // global input row
      if (m >= 256 || ox >= nf) continue;
      const int py = oyl & 1, px = ox & 1, n = ox >> 1;
      float acc = bs;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int kh = py + 2 * a;
        if (kh > 4) continue;
        const int rr = mr + ((py + 2 - kh) >> 1);                        // slab row (halo rows are zero-filled by TMA)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int kw = px + 2 * c;
          if (kw > 4) continue;
          const int cc = n + ((px + 2 - kw) >> 1);
          if (cc < 0 || cc >= 64) continue;
          acc += P[(kh * 5 + kw) * kD6PPitch + rr * 64 + cc];
        }
      }
      float mval = 1.0f / (1.0f + __expf(-acc));                         // torch.sigmoid, model.py:200
      if (flags & SVS_FLAG_INVERT) mval = 1.0f - mval;                   // inference.py:102
      const int oy = 2 * m + py;
      out_b[oy * out_sf + ox * out_st] = mval * mixv[j];                 // inference.py:107 (mixv = 1 without APPLY_MASK)
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host

__global__ void d6_pack_weights_kernel(const float* __restrict__ w_fold /*[25][32]*/, int tf32, void* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;                 // 32 rows (taps) x 32 k (channels)
  if (i >= 32 * 32) return;
  const int tap = i >> 5;
  const float v = tap < kD6Taps ? w_fold[i] : 0.0f;
  if (tf32) static_cast<float*>(out)[i] = round_tf32(v);
  else static_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
}

int d6_plan(svs_unet_plan* plan, cudaStream_t st) {
  const bool tf32 = plan->precision == SVS_PRECISION_TF32;
  const int es = plan->elem_size;
  SVS_CUDA_TRY(cudaMalloc(&plan->d6_weights, 32 * 32 * es));
  d6_pack_weights_kernel<<<4, 256, 0, st>>>(plan->w_fold[11], tf32 ? 1 : 0, plan->d6_weights);
  SVS_CHECK_LAUNCH("d6_pack_weights_kernel");
  const cuuint64_t dims[2] = {32, 32};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(32 * es)};
  const cuuint32_t box[2] = {32, 32};
  int rc = encode_tensor_map(&plan->d6_tmap_w, tf32, 2, plan->d6_weights, dims, strides, box, 32 * es);
  if (rc != SVS_OK) return rc;
  plan->d6_enabled = true;
  return SVS_OK;
}

void d6_free(svs_unet_plan* plan) {
  if (plan->d6_weights) cudaFree(plan->d6_weights);
  plan->d6_weights = nullptr;
  plan->d6_enabled = false;
}

// (re)pack the [25][32] folded weights into the taps-as-N operand (training repacks every step)
int d6_pack(const float* w_fold, bool tf32, void* d6_weights, cudaStream_t st) {
  d6_pack_weights_kernel<<<4, 256, 0, st>>>(w_fold, tf32 ? 1 : 0, d6_weights);
  SVS_CHECK_LAUNCH("d6_pack_weights_kernel");
  return SVS_OK;
}

int d6_launch_raw(bool tf32, const CUtensorMap& tmap_w, const float* bias, const void* cat1, const svs_patch_view* in,
                  const svs_patch_view* out, const int32_t* in_frames, int batch, int flags, cudaStream_t st) {
  const int es = tf32 ? 4 : 2;
  CUtensorMap ta;
  const cuuint64_t dims[4] = {32, 64, 256, static_cast<cuuint64_t>(batch)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(32 * es), static_cast<cuuint64_t>(64 * 32 * es),
                                 static_cast<cuuint64_t>(256) * 64 * 32 * es};
  const cuuint32_t box[4] = {32, 64, kD6Rows, 1};
  int rc = encode_tensor_map(&ta, tf32, 4, const_cast<void*>(cat1), dims, strides, box, 32 * es);
  if (rc != SVS_OK) return rc;
  dim3 grid((256 + kD6Interior - 1) / kD6Interior, batch);
  auto dense_view = [](const svs_patch_view* v) {
    return v->patch_off == nullptr && v->stride_t == 1 && v->stride_f == SVS_PATCH_FRAMES &&
           v->stride_b == static_cast<int64_t>(SVS_PATCH_BINS) * SVS_PATCH_FRAMES &&
           (reinterpret_cast<uintptr_t>(v->base) & 7) == 0;
  };
  const bool dense = dense_view(in) && dense_view(out);
  auto launch = [&](auto kern, size_t smem) -> int {
    SVS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    SVS_CUDA_TRY(launch_pdl(kern, grid, dim3(kD6Threads), smem, st, ta, tmap_w, bias,
                            static_cast<const float*>(in->base), in->patch_off, in->stride_b, in->stride_f,
                            in->stride_t, out->base, out->patch_off, out->stride_b, out->stride_f, out->stride_t,
                            in_frames, flags));
    return SVS_OK;
  };
  if (tf32) return dense ? launch(deconv6_tc_kernel<true, 128, true>, d6_smem_bytes<128>())
                         : launch(deconv6_tc_kernel<true, 128, false>, d6_smem_bytes<128>());
  return dense ? launch(deconv6_tc_kernel<false, 64, true>, d6_smem_bytes<64>())
               : launch(deconv6_tc_kernel<false, 64, false>, d6_smem_bytes<64>());
}

int d6_make_weight_map(void* d6_weights, bool tf32, CUtensorMap* out) {
  const int es = tf32 ? 4 : 2;
  const cuuint64_t dims[2] = {32, 32};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(32 * es)};
  const cuuint32_t box[2] = {32, 32};
  return encode_tensor_map(out, tf32, 2, d6_weights, dims, strides, box, 32 * es);
}

int d6_launch(const svs_unet_plan* plan, const Workspace& ws, const svs_patch_view* in, const svs_patch_view* out,
              const int32_t* in_frames, int batch, int flags, cudaStream_t st) {
  return d6_launch_raw(plan->precision == SVS_PRECISION_TF32, plan->d6_tmap_w,
                       static_cast<const float*>(plan->b_fold[11]), ws.buf[BUF_CAT1], in, out, in_frames, batch, flags, st);
}

}  // namespace svs
