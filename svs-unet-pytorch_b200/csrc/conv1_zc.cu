// U1 (conv1 + folded BatchNorm + LeakyReLU, reference model.py:47-51,176) on tcgen05 WITHOUT building im2col rows.
//
// conv1 has ONE input channel: an implicit-GEMM row would be 25 taps gathered from five image rows, and building
// those rows in shared memory made conv1_tc.cu L1/shared-bandwidth bound (72 % l1tex, 2.7 TB/s of HBM).  Here the
// GEMM is turned so that the image itself is the A operand, untouched, in fp32 (kind::tf32 reads fp32 bits):
//
//     for kernel row kh:   D[y][(xl, co)] += sum_j  X[2y + kh - 2][16 bx - 4 + j] * Wb[kh][(xl, co)][j]
//
//   M = 128 output rows y, N = 8 output columns x 16 channels, K = a window of 32 input columns, and
//   Wb[kh][(xl, co)][j] = w[co][kh][kw] where j = 2 + 2 xl + kw (banded: the stride-2 horizontal taps), else 0;
//   the window starts at column 16 bx - 4 because TMA wants the innermost start 16-byte aligned.
//
// With the patch viewed as (t, row parity, row / 2, patch), the rows 2y + kh - 2 for 128 consecutive y are 128
// consecutive rows of ONE parity plane, so two TMA boxes per unit (130 x 128 B each, zero-filled outside the
// image = the conv padding) are the A operand of all five kernel rows: kh selects the plane and a start-address
// shift of 0 / 1 / 2 rows (128 B) of the SWIZZLE_128B tile, exactly like the halo slabs of zc_conv.cu.  The banded
// weights (5 x 16 KB, TF32) stay resident in shared memory; window columns 24..31 carry only zeros, so 3 of the 4
// K steps are issued: 15 MMAs (128 x 128 x 8) per 1024 output pixels.  Shared-memory traffic per unit is the two
// TMA writes plus the MMA reads -- no LDS/STS by threads at all; what remains is the HBM stream.
#include "tc_conv_common.cuh"

namespace svs {

constexpr int kC1zThreads = 192;                 // warp 0: TMA, warp 1: MMA issue, warps 2..5: epilogue
constexpr int kC1zStages = 3;
constexpr int kC1zSlabRows = 130;
constexpr int kC1zSlabBytes = kC1zSlabRows * 128;            // 16640 delivered per plane
constexpr int kC1zSlabSlot = 17 * 1024;                      // 1024-aligned slot
constexpr int kC1zStageBytes = 2 * kC1zSlabSlot;
constexpr int kC1zWBytes = 5 * 128 * 128;                    // banded weights: 5 x [128 n][32 k] fp32
// bf16 epilogue staging: each epilogue warp transposes its 32 rows x (8 px x 16 ch) through a private buffer so a
// store instruction covers 2 rows x 256 B (8 cache lines) instead of 32 rows x 16 B (32 lines): the row-per-thread
// stores cost one L1 tag cycle per line and were slower than the MMAs
constexpr int kC1zOutPitch = 256 + 16;                       // bytes per staged row (+16: 4-wavefront STS)
constexpr int kC1zOutBytes = 4 * 32 * kC1zOutPitch;
constexpr size_t kC1zSmemBytes = static_cast<size_t>(kC1zStages) * kC1zStageBytes + kC1zWBytes + kC1zOutBytes + 1024 + 256;

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <typename OutT>
__global__ void __launch_bounds__(kC1zThreads)
conv1_zc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const float* __restrict__ bias, OutT* __restrict__ out, int out_pitch, int out_coff, int n_units) {
  constexpr bool kTf32Out = sizeof(OutT) == 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w_base = smem_base + kC1zStages * kC1zStageBytes;
  const uint32_t stage_out = w_base + kC1zWBytes;
  const uint32_t bar_base = stage_out + kC1zOutBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kC1zStages + s); };
  auto tmem_full = [&](int a) { return bar_base + 8u * (2 * kC1zStages + a); };
  auto tmem_empty = [&](int a) { return bar_base + 8u * (2 * kC1zStages + 2 + a); };
  const uint32_t w_bar = bar_base + 8u * (2 * kC1zStages + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kC1zStages + 5);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kC1zStages * kC1zStageBytes + kC1zWBytes + kC1zOutBytes +
                                           8 * (2 * kC1zStages + 5));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kC1zStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), 4); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  // unit u = ((patch * 2 + row half) * 8 + column block): 128 output rows x 8 output columns x 16 channels
  if (warp == 0) {
    if (elect_one_sync()) {                       // weights do not depend on the previous kernel
      mbar_expect_tx(w_bar, kC1zWBytes);
      for (int kh = 0; kh < 5; ++kh) tma_load_2d(w_base + kh * 128 * 128, &tmap_w, w_bar, 0, kh * 128);
    }
    __syncwarp();
    pdl_wait();
    int it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
      const int bx = u & 7, mt = (u >> 3) & 1, b = u >> 4;
      const int s = it % kC1zStages;
      mbar_wait(empty_bar(s), ((it / kC1zStages) & 1) ^ 1);
      if (elect_one_sync()) {
        const uint32_t dst = smem_base + s * kC1zStageBytes;
        mbar_expect_tx(full_bar(s), 2 * kC1zSlabBytes);
        // rows (2 yh + parity) for yh = 128 mt - 1 .. 128 mt + 128, window columns 16 bx - 4 .. 16 bx + 27
        tma_load_4d(dst, &tmap_x, full_bar(s), 16 * bx - 4, 0, 128 * mt - 1, b);
        tma_load_4d(dst + kC1zSlabSlot, &tmap_x, full_bar(s), 16 * bx - 4, 1, 128 * mt - 1, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc<true, 128>();
    mbar_wait(w_bar, 0);
    tc_fence_after();
    int it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
      const int s = it % kC1zStages;
      const int as = it & 1;
      mbar_wait(tmem_empty(as), ((it >> 1) & 1) ^ 1);
      mbar_wait(full_bar(s), (it / kC1zStages) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * 128;
      dispatch_stage<0, kC1zStages>(s, [&](auto sc) {
        constexpr int S = decltype(sc)::value;
        const uint32_t a0 = smem_base + S * kC1zStageBytes;
        if (elect_one_sync()) {
#pragma unroll
          for (int kh = 0; kh < 5; ++kh) {
            // kh = 0, 2, 4: even plane, rows y - 1, y, y + 1;  kh = 1, 3: odd plane, rows y - 1, y
            const uint64_t da = make_smem_desc<128>(a0 + (kh & 1) * kC1zSlabSlot + (kh >> 1) * 128);
            const uint64_t db = make_smem_desc<128>(w_base + kh * 128 * 128);
#pragma unroll
            for (int k = 0; k < 3; ++k)                     // window columns 24..31 only meet zero weights
              umma<true>(tmem_d, da + 2u * k, db + 2u * k, idesc, (kh > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(bar_base + 8u * (kC1zStages + S));     // empty_bar(S)
          umma_commit(tmem_full(as));
        }
        __syncwarp();
      });
    }
  } else {
    const int q = warp & 3;
    const int r = 32 * q + lane;                             // accumulator lane = output row within the unit
    float bv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bv[i] = __ldg(&bias[i]);
    pdl_wait();                                              // before the first global store
    int it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
      const int bx = u & 7, mt = (u >> 3) & 1, b = u >> 4;
      const int as = it & 1;
      mbar_wait(tmem_full(as), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * 128;
      if constexpr (kTf32Out) {
        const int y = 128 * mt + r;
        OutT* dst = out + (static_cast<size_t>(b * 256 + y) * 64 + 8 * bx) * out_pitch + out_coff;
#pragma unroll 2
        for (int c = 0; c < 128; c += 32) {                  // 2 output pixels x 16 channels per step
          uint32_t v[32];
          tmem_ld16(taddr + c, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          tmem_ld16(taddr + c + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float t = __uint_as_float(v[16 * h + i]) + bv[i];
              f[i] = round_tf32(fmaxf(t, 0.2f * t));          // LeakyReLU(0.2), model.py:50
            }
            store16(dst + static_cast<size_t>((c >> 4) + h) * out_pitch, f);
          }
        }
      } else {
        const uint32_t mine = stage_out + q * 32 * kC1zOutPitch;
#pragma unroll 2
        for (int c = 0; c < 128; c += 32) {
          uint32_t v[32];
          tmem_ld16(taddr + c, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          tmem_ld16(taddr + c + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float t0 = __uint_as_float(v[16 * h + 2 * i]) + bv[2 * i];
              const float t1 = __uint_as_float(v[16 * h + 2 * i + 1]) + bv[2 * i + 1];
              const __nv_bfloat162 pk = __floats2bfloat162_rn(fmaxf(t0, 0.2f * t0), fmaxf(t1, 0.2f * t1));
              w[i] = *reinterpret_cast<const uint32_t*>(&pk);
            }
            const uint32_t a = mine + lane * kC1zOutPitch + ((c >> 4) + h) * 32;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 16), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
          }
        }
        __syncwarp();
        // lanes 0..15: the 8 pixels x 32 B of row 2i, lanes 16..31: row 2i + 1
        const int half = lane >> 4, chunk = lane & 15;
        char* gbase = reinterpret_cast<char*>(out + (static_cast<size_t>(b * 256 + 128 * mt + 32 * q) * 64 + 8 * bx) * out_pitch + out_coff) +
                      (chunk >> 1) * (out_pitch * 2) + (chunk & 1) * 16;
        const size_t row_bytes = static_cast<size_t>(64) * out_pitch * 2;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int row = 2 * i + half;
          uint4 val;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                       : "r"(mine + row * kC1zOutPitch + chunk * 16)
                       : "memory");
          *reinterpret_cast<uint4*>(gbase + row * row_bytes) = val;
        }
        __syncwarp();                                        // staging buffer is reused by the next unit
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty(as));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// Wb[kh][n = xl * 16 + co][j] = w[kh][kw = j - 2 - 2 xl][co] (tap-major BN-folded weights [25][1][16]), TF32-rounded
__global__ void c1z_pack_weights_kernel(const float* __restrict__ w_fold, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 5 * 128 * 32) return;
  const int j = i & 31, n = (i >> 5) & 127, kh = i >> 12;
  const int xl = n >> 4, co = n & 15;
  const int kw = j - 2 - 2 * xl;
  out[i] = (kw >= 0 && kw < 5) ? round_tf32(w_fold[(kh * 5 + kw) * 16 + co]) : 0.0f;
}

int c1z_plan(svs_unet_plan* plan, cudaStream_t st) {
  static const bool off = [] { const char* e = std::getenv("SVS_C1Z_DISABLE"); return e && e[0] == '1'; }();
  if (off || plan->precision == SVS_PRECISION_FP32) return SVS_OK;
  SVS_CUDA_TRY(cudaMalloc(&plan->c1z_weights, kC1zWBytes));
  c1z_pack_weights_kernel<<<(5 * 128 * 32 + 255) / 256, 256, 0, st>>>(plan->w_fold[0], plan->c1z_weights);
  SVS_CHECK_LAUNCH("c1z_pack_weights_kernel");
  const cuuint64_t dims[2] = {32, 5 * 128};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {32, 128};
  int rc = encode_tensor_map(&plan->c1z_tmap_w, true, 2, plan->c1z_weights, dims, strides, box, 128);
  if (rc != SVS_OK) return rc;
  plan->c1z_enabled = true;
  return SVS_OK;
}

void c1z_free(svs_unet_plan* plan) {
  if (plan->c1z_weights) cudaFree(plan->c1z_weights);
  plan->c1z_weights = nullptr;
  plan->c1z_enabled = false;
}

int c1z_launch(const svs_unet_plan* plan, const Workspace& ws, const svs_patch_view* in, int batch, cudaStream_t st) {
  CUtensorMap tx;
  {
    // dense fp32 patches [b][512][128] viewed as (t, row parity, row / 2, b)
    const cuuint64_t dims[4] = {128, 2, 256, static_cast<cuuint64_t>(batch)};
    const cuuint64_t strides[3] = {128 * 4, 2 * 128 * 4, 512 * 128 * 4};
    const cuuint32_t box[4] = {32, 1, kC1zSlabRows, 1};
    int rc = encode_tensor_map(&tx, true, 4, in->base, dims, strides, box, 128);
    if (rc != SVS_OK) return rc;
  }
  const LayerGeom& g = kLayers[0];
  const int n_units = batch * 16;
  int grid = num_sms();
  if (grid > n_units) grid = n_units;
  if (plan->precision == SVS_PRECISION_TF32) {
    SVS_CUDA_TRY(cudaFuncSetAttribute(conv1_zc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(kC1zSmemBytes)));
    SVS_CUDA_TRY(launch_pdl(conv1_zc_kernel<float>, dim3(grid), dim3(kC1zThreads), kC1zSmemBytes, st, tx,
                            plan->c1z_tmap_w, static_cast<const float*>(plan->b_fold[0]),
                            reinterpret_cast<float*>(ws.buf[g.out_buf]), kBufGeom[g.out_buf].c, g.out_coff, n_units));
  } else {
    SVS_CUDA_TRY(cudaFuncSetAttribute(conv1_zc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(kC1zSmemBytes)));
    SVS_CUDA_TRY(launch_pdl(conv1_zc_kernel<__nv_bfloat16>, dim3(grid), dim3(kC1zThreads), kC1zSmemBytes, st, tx,
                            plan->c1z_tmap_w, static_cast<const float*>(plan->b_fold[0]),
                            reinterpret_cast<__nv_bfloat16*>(ws.buf[g.out_buf]), kBufGeom[g.out_buf].c, g.out_coff,
                            n_units));
  }
  return SVS_OK;
}

}  // namespace svs
