// U1 on the tensor cores: Conv2d(1 -> 16, 5x5, stride 2, pad 2) + folded BatchNorm + LeakyReLU(0.2)
// (reference model.py:47-51,176) for DENSE patch batches (B,1,512,128).
//
// With one input channel the implicit GEMM has K = 25 taps (padded to 32) and N = 16: far too thin to
// stage through TMA boxes, and on the CUDA cores the layer is bound by its 400 FMA per output pixel
// (conv1_kernel: 43 us per 64 patches against a 8 us HBM bound).  Here the CTA
//   1. lands the 19 x 136 fp32 input window of 8 output rows with ONE TMA box (conv zero padding = OOB fill),
//   2. builds the im2col operand itself: each thread converts the 25 taps of its pixel and writes one
//      K-major row (32 elements) straight into the swizzled layout the UMMA descriptor reads
//      (generic-proxy stores + fence.proxy.async),
//   3. issues 4 M-tiles x (K = 32) tcgen05.mma against the resident 16 x 32 weight tile,
//   4. drains TMEM -> bias + LeakyReLU -> bf16/fp32 -> the skip half of concat buffer 1.
// Several CTAs per SM overlap each other's phases.  Strided / ragged patch views (the fused song
// pipeline) keep using conv1_kernel.
#include "unet_internal.cuh"
#include "tc_ptx.cuh"

namespace svs {

constexpr int kC1tRows = 8;                       // output rows per CTA = 4 M-tiles of 2 rows x 64 cols
constexpr int kC1tInRows = 2 * kC1tRows + 3;      // 19
constexpr int kC1tPitch = 136;                    // 4 + 128 + 4 floats, box starts at t = -4
constexpr int kC1tThreads = 160;                  // warps 0..3: im2col + epilogue, warp 4: TMA + MMA

template <bool kTf32>
constexpr size_t c1t_smem_bytes() {
  return sizeof(float) * kC1tInRows * kC1tPitch + 1024 /*align*/ + 4 * 128 * (kTf32 ? 128 : 64) + 2048 /*weights*/ + 256;
}

template <typename OutT, bool kTf32>
__global__ void __launch_bounds__(kC1tThreads)
conv1_tc_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
                const float* __restrict__ bias, OutT* __restrict__ out, int out_pitch, int out_coff) {
  constexpr int kRowBytes = kTf32 ? 128 : 64;     // one K-major operand row: 32 fp32 / 32 bf16
  constexpr int kSwz = kRowBytes;
  constexpr int kATile = 128 * kRowBytes;
  constexpr int kKSteps = kRowBytes / 32;
  constexpr int kInBytes = sizeof(float) * kC1tInRows * kC1tPitch;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: [A tiles 4 x kATile][W 16 x kRowBytes (1024-aligned)][input window][barriers]
  const uint32_t a_base = smem_base;
  const uint32_t w_base = a_base + 4 * kATile;
  constexpr int kWBytes = 16 * kRowBytes;
  const uint32_t in_base = w_base + 2048;
  const float* tile = reinterpret_cast<const float*>(smem_gen + 4 * kATile + 2048);
  const uint32_t bar_in = in_base + ((kInBytes + 15) / 16) * 16;
  const uint32_t bar_mma = bar_in + 8;
  const uint32_t tmem_slot = bar_in + 16;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + 4 * kATile + 2048 + ((kInBytes + 15) / 16) * 16 + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int oh0 = blockIdx.x * kC1tRows;

  if (threadIdx.x == 0) {
    mbar_init(bar_in, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_in);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 4) tmem_alloc<64>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();                                     // the input may come from a preceding kernel; cat1 is reused across forwards

  if (warp == 4) {
    if (elect_one_sync()) {
      mbar_expect_tx(bar_in, kInBytes + kWBytes);
      // input window: t in [-4, 132), f in [2 oh0 - 2, 2 oh0 + 17), patch b; OOB -> 0 (conv padding)
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
          " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(in_base),
          "l"(reinterpret_cast<uint64_t>(&tmap_in)), "r"(bar_in), "r"(-4), "r"(2 * oh0 - 2), "r"(b)
          : "memory");
      tma_load_2d(w_base, &tmap_w, bar_in, 0, 0);
    }
    __syncwarp();
  } else {
    // ---- im2col: thread w owns operand row w of each of the 4 M-tiles (pixel = row 2 mt + w/64, col w%64) ----
    const int w = threadIdx.x;
    const int ox = w & 63, rsel = w >> 6;
    mbar_wait(bar_in, 0);
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int oyl = 2 * mt + rsel;
      float taps[32];
#pragma unroll
      for (int kh = 0; kh < 5; ++kh) {
        // input columns t = 2 ox - 2 .. 2 ox + 2  ->  window columns 2 ox + 2 .. 2 ox + 6
        const float* row = tile + (2 * oyl + kh) * kC1tPitch + 2 * ox + 2;
        const float2 p0 = *reinterpret_cast<const float2*>(row);
        const float2 p1 = *reinterpret_cast<const float2*>(row + 2);
        taps[kh * 5 + 0] = p0.x; taps[kh * 5 + 1] = p0.y; taps[kh * 5 + 2] = p1.x; taps[kh * 5 + 3] = p1.y;
        taps[kh * 5 + 4] = row[4];
      }
#pragma unroll
      for (int i = 25; i < 32; ++i) taps[i] = 0.0f;
      const uint32_t row_addr = a_base + mt * kATile + w * kRowBytes;
      if constexpr (kTf32) {
#pragma unroll
        for (int i = 0; i < 25; ++i) taps[i] = round_tf32(taps[i]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {                                   // 8 x 16-byte chunks, Swizzle<3,4,3>
          const uint32_t dst = row_addr + (((c ^ (w & 7)) & 7) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(taps[4 * c]), "f"(taps[4 * c + 1]),
                       "f"(taps[4 * c + 2]), "f"(taps[4 * c + 3])
                       : "memory");
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {                                   // 4 x 16-byte chunks, Swizzle<2,4,3>
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(taps[8 * c + 2 * j], taps[8 * c + 2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          const uint32_t dst = row_addr + (((c ^ ((w >> 1) & 3)) & 3) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                       : "memory");
        }
      }
    }
    fence_proxy_async();                          // generic-proxy stores -> visible to the tensor core (async proxy)
  }
  __syncthreads();
  if (warp == 4) {
    mbar_wait(bar_in, 0);                         // weights landed
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc<kTf32, 16>();
    const uint64_t dw = make_smem_desc<kSwz>(w_base);
    if (elect_one_sync()) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const uint64_t da = make_smem_desc<kSwz>(a_base + mt * kATile);
#pragma unroll
        for (int k = 0; k < kKSteps; ++k) umma<kTf32>(tmem_base + mt * 16, da + 2u * k, dw + 2u * k, idesc, k > 0 ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    __syncwarp();
  } else {
    // ---- epilogue: TMEM lane w of each accumulator -> bias + LeakyReLU -> concat buffer (skip half) ----
    const int w = threadIdx.x;
    const int ox = w & 63, rsel = w >> 6;
    float bv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bv[i] = __ldg(&bias[i]);
    mbar_wait(bar_mma, 0);
    tc_fence_after();
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      uint32_t v[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(32 * warp) << 16) + mt * 16, v);
      tmem_ld_wait();
      const int oy = oh0 + 2 * mt + rsel;
      OutT* dst = out + ((static_cast<size_t>(b) * 256 + oy) * 64 + ox) * out_pitch + out_coff;
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float x = __uint_as_float(v[i]) + bv[i];
        f[i] = x > 0.0f ? x : 0.2f * x;
        if constexpr (kTf32) f[i] = round_tf32(f[i]);          // consumed by a kind::tf32 MMA
      }
      if constexpr (sizeof(OutT) == 2) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
          pk[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        reinterpret_cast<uint4*>(dst)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        reinterpret_cast<uint4*>(dst)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          reinterpret_cast<float4*>(dst)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void c1t_pack_weights_kernel(const float* __restrict__ w_fold /*[25][1][16]*/, int tf32, void* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;                 // [16 co][32 k]
  if (i >= 16 * 32) return;
  const int co = i >> 5, k = i & 31;
  const float v = k < 25 ? w_fold[k * 16 + co] : 0.0f;
  if (tf32) static_cast<float*>(out)[i] = round_tf32(v);
  else static_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
}

int c1t_plan(svs_unet_plan* plan, cudaStream_t st) {
  const bool tf32 = plan->precision == SVS_PRECISION_TF32;
  const int es = plan->elem_size;
  SVS_CUDA_TRY(cudaMalloc(&plan->c1_weights, 16 * 32 * es));
  c1t_pack_weights_kernel<<<2, 256, 0, st>>>(plan->w_fold[0], tf32 ? 1 : 0, plan->c1_weights);
  SVS_CHECK_LAUNCH("c1t_pack_weights_kernel");
  const cuuint64_t dims[2] = {32, 16};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(32 * es)};
  const cuuint32_t box[2] = {32, 16};
  int rc = encode_tensor_map(&plan->c1_tmap_w, tf32, 2, plan->c1_weights, dims, strides, box, 32 * es);
  if (rc != SVS_OK) return rc;
  plan->c1_enabled = true;
  return SVS_OK;
}

void c1t_free(svs_unet_plan* plan) {
  if (plan->c1_weights) cudaFree(plan->c1_weights);
  plan->c1_weights = nullptr;
  plan->c1_enabled = false;
}

// dense, unpadded patch batches only
bool c1t_applicable(const svs_patch_view* in, const int32_t* in_frames) {
  return in->patch_off == nullptr && in_frames == nullptr && in->stride_t == 1 && in->stride_f == SVS_PATCH_FRAMES &&
         in->stride_b == static_cast<int64_t>(SVS_PATCH_BINS) * SVS_PATCH_FRAMES &&
         (reinterpret_cast<uintptr_t>(in->base) & 15) == 0;
}

int c1t_launch(const svs_unet_plan* plan, const Workspace& ws, const svs_patch_view* in, int batch, cudaStream_t st) {
  const bool tf32 = plan->precision == SVS_PRECISION_TF32;
  CUtensorMap tin;
  {
    // fp32 input (t, f, b); the fp32 data type is used for both precisions (conversion happens in the kernel)
    const cuuint64_t dims[3] = {128, 512, static_cast<cuuint64_t>(batch)};
    const cuuint64_t strides[2] = {128 * 4, 512 * 128 * 4};
    const cuuint32_t box[3] = {kC1tPitch, kC1tInRows, 1};
    int rc = encode_tensor_map(&tin, true, 3, in->base, dims, strides, box, 0);
    if (rc != SVS_OK) return rc;
  }
  const LayerGeom& g = kLayers[0];
  dim3 grid(256 / kC1tRows, batch);
  if (tf32) {
    constexpr size_t smem = c1t_smem_bytes<true>();
    SVS_CUDA_TRY(cudaFuncSetAttribute(conv1_tc_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    SVS_CUDA_TRY(launch_pdl(conv1_tc_kernel<float, true>, grid, dim3(kC1tThreads), smem, st, tin, plan->c1_tmap_w,
                            static_cast<const float*>(plan->b_fold[0]), reinterpret_cast<float*>(ws.buf[g.out_buf]),
                            kBufGeom[g.out_buf].c, g.out_coff));
  } else {
    constexpr size_t smem = c1t_smem_bytes<false>();
    SVS_CUDA_TRY(cudaFuncSetAttribute(conv1_tc_kernel<__nv_bfloat16, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    SVS_CUDA_TRY(launch_pdl(conv1_tc_kernel<__nv_bfloat16, false>, grid, dim3(kC1tThreads), smem, st, tin,
                            plan->c1_tmap_w, static_cast<const float*>(plan->b_fold[0]),
                            reinterpret_cast<__nv_bfloat16*>(ws.buf[g.out_buf]), kBufGeom[g.out_buf].c, g.out_coff));
  }
  return SVS_OK;
}

}  // namespace svs
