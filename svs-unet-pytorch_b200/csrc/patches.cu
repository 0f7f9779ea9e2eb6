// Patch staging between the frame-major song spectrogram ([frames][513], librosa's F-ordered (513, T)) and the dense
// [n][512][128] patch batches of the UNet (reference inference.py:74-97 builds them one patch at a time with a
// C-order copy of an F-order slice; inference.py:110-127 undoes it).
//
// svs_unet_forward can read / write the spectrogram directly through a strided patch view, but conv1 and
// deconv6 are then off their fast paths (the 2052-byte frame pitch is not a legal TMA stride, and the mixture /
// output accesses of deconv6 are 4-byte gathers): 88 us instead of 44 us per 64 patches.  A tiled transpose at
// HBM speed (2 x 16.8 MB per 64 patches, ~6 us each way) with the per-song normalisation (data.py:85,105) folded
// into the gather is cheaper, and it replaces the separate normalise pass and the zero fill of the output.
#include "svs_common.cuh"

namespace svs {

constexpr int kPatchF = 512, kPatchT = 128, kTileF = 32;

// grid (4, n): block = 32 consecutive frames x all 513 bins of one patch.  The 32 frame rows are one contiguous
// 65.7 KB run of the spectrogram, so the reads are perfectly coalesced (a 32-frequency tile read 128-byte pieces of
// 2,052-byte rows: 5 sectors per 4 used); the transposed write-out is one 128-byte line per warp store.
constexpr int kGatherT = 32;
constexpr int kGatherPitch = SVS_N_BINS;                       // 513 = 1 mod 32: column reads are conflict free
constexpr size_t kGatherSmemBytes = sizeof(float) * kGatherT * kGatherPitch;

__global__ void __launch_bounds__(256)
patches_gather_kernel(const float* __restrict__ spec, const int64_t* __restrict__ patch_off,
                      const int32_t* __restrict__ in_frames, const float* __restrict__ norm,
                      float* __restrict__ patches) {
  extern __shared__ float tile[];                              // [32 frames][513 bins]
  const int p = blockIdx.y, t0 = blockIdx.x * kGatherT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int valid = in_frames ? in_frames[p] : kPatchT;
  float nrm = norm ? norm[p] : 1.0f;
  if (nrm == 0.0f) nrm = 1.0f;                                  // reference data.py:85
  // patch_off points at bin 1 of the patch's first frame; row t of the tile starts at that frame's DC bin
  const float* src = spec + patch_off[p] - 1 + static_cast<int64_t>(t0) * SVS_N_BINS;
  const int rows = min(kGatherT, max(0, valid - t0));           // frames >= valid are zero padding (inference.py:90-92)
  for (int i = threadIdx.x; i < rows * SVS_N_BINS; i += 256) tile[i] = __ldg(src + i);
  __syncthreads();
  float* dst = patches + static_cast<int64_t>(p) * kPatchF * kPatchT + t0 + lane;
#pragma unroll 4
  for (int f = warp; f < kPatchF; f += 8)
    dst[f * kPatchT] = lane < rows ? tile[lane * kGatherPitch + f + 1] / nrm : 0.0f;
}

__global__ void __launch_bounds__(256)
patches_scatter_kernel(const float* __restrict__ patches, const int64_t* __restrict__ patch_off,
                       const int32_t* __restrict__ in_frames, float* __restrict__ spec, int dc_zero) {
  __shared__ float tile[kPatchT][kTileF + 1];
  const int p = blockIdx.y, f0 = blockIdx.x * kTileF;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int valid = in_frames ? in_frames[p] : kPatchT;
  const float* src = patches + (static_cast<int64_t>(p) * kPatchF + f0) * kPatchT;
#pragma unroll 4
  for (int i = ty; i < kTileF * 4; i += 8) {
    const int f = i >> 2, t = 32 * (i & 3) + tx;
    tile[t][f] = __ldg(src + f * kPatchT + t);
  }
  __syncthreads();
  float* dst = spec + patch_off[p] + f0 + tx;
#pragma unroll 4
  for (int t = ty; t < valid; t += 8) dst[static_cast<int64_t>(t) * SVS_N_BINS] = tile[t][tx];      // crop: inference.py:113
  if (dc_zero && blockIdx.x == 0)                                                                  // inference.py:123
    for (int t = threadIdx.x; t < valid; t += 256) spec[patch_off[p] - 1 + static_cast<int64_t>(t) * SVS_N_BINS] = 0.0f;
}

}  // namespace svs

extern "C" int svs_patches_gather(const float* spec, const int64_t* patch_off, const int32_t* in_frames,
                                  const float* norm, float* patches, int n, void* stream) {
  using namespace svs;
  SVS_REQUIRE(spec && patch_off && patches, "svs_patches_gather: null pointer");
  SVS_REQUIRE(n >= 0 && n <= 65535, "svs_patches_gather: n must be in [0, 65535]");
  if (n == 0) return SVS_OK;
  SVS_CUDA_TRY(cudaFuncSetAttribute(patches_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(kGatherSmemBytes)));
  patches_gather_kernel<<<dim3(kPatchT / kGatherT, n), 256, kGatherSmemBytes, static_cast<cudaStream_t>(stream)>>>(
      spec, patch_off, in_frames, norm, patches);
  SVS_CHECK_LAUNCH("patches_gather_kernel");
  return SVS_OK;
}

extern "C" int svs_patches_scatter(const float* patches, const int64_t* patch_off, const int32_t* in_frames,
                                   float* spec, int n, int dc_zero, void* stream) {
  using namespace svs;
  SVS_REQUIRE(spec && patch_off && patches, "svs_patches_scatter: null pointer");
  SVS_REQUIRE(n >= 0 && n <= 65535, "svs_patches_scatter: n must be in [0, 65535]");
  if (n == 0) return SVS_OK;
  patches_scatter_kernel<<<dim3(kPatchF / kTileF, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      patches, patch_off, in_frames, spec, dc_zero);
  SVS_CHECK_LAUNCH("patches_scatter_kernel");
  return SVS_OK;
}
