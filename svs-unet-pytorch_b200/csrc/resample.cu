// K0: decode-side resampler — PCM -> mono -> polyphase rational resampling to the model's 8192 Hz on the GPU.
//
// Replaces the `librosa.load(path, sr=8192, mono=True)` of reference data.py:78,94 AFTER the file bytes are read:
// int16 -> float (libsndfile: / 32768), channel mean (librosa.to_mono), then a windowed-sinc polyphase FIR
//     y[n] = sum_j h[(n + r) * down - pre - j * up] * x[j]
// with the filter / padding conventions of scipy.signal.resample_poly (Kaiser beta 5, half length 10 * max(up,
// down), gain `up`, n_out = ceil(n_in * up / down)); the host side designs h and passes it in POLYPHASE order
// hp[phase][tap] = h[phase + tap * up].  librosa's default soxr_hq resampler is a different (unpublished-coefficient)
// filter: parity with the reference is therefore at the level of "a high-quality band-limited resampler", and the
// oracle for this kernel is scipy.signal.resample_poly in float64 (oracle/resample_oracle.py).
//
// One thread per output sample, ~len(h)/up taps each (108 for 44100 -> 8192); the 0.9 MB filter table lives in L2,
// the input is read once from HBM (4 bytes per stereo int16 frame).  Ragged batch of songs like K1 / K2.
#include "svs_common.cuh"

namespace svs {

template <typename InT>
__device__ __forceinline__ float load_mono(const InT* __restrict__ x, int64_t frame, int channels) {
  float acc = 0.0f;
  const InT* p = x + frame * channels;
  for (int c = 0; c < channels; ++c) {
    if constexpr (sizeof(InT) == 2) acc += static_cast<float>(__ldg(p + c)) * (1.0f / 32768.0f);
    else acc += __ldg(p + c);
  }
  return channels == 1 ? acc : acc / static_cast<float>(channels);
}

template <typename InT>
__global__ void __launch_bounds__(256)
resample_poly_kernel(const InT* __restrict__ in, int channels, const int64_t* __restrict__ in_off,
                     const int64_t* __restrict__ out_off, int up, int down, int64_t pre_pad, int64_t pre_remove,
                     const float* __restrict__ hp, int taps, float* __restrict__ out) {
  const int song = blockIdx.y;
  const int64_t i0 = in_off[song], n_in = in_off[song + 1] - i0;
  const int64_t o0 = out_off[song], n_out = out_off[song + 1] - o0;
  const InT* x = in + i0 * channels;
  for (int64_t n = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; n < n_out;
       n += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t q = (n + pre_remove) * down - pre_pad;        // index into h of the tap that meets x[q / up]
    int64_t j_hi = q >= 0 ? q / up : -((-q + up - 1) / up);     // floor(q / up)
    const int ph = static_cast<int>(q - j_hi * up);
    const float* h = hp + static_cast<size_t>(ph) * taps;
    float acc = 0.0f;
    for (int t = 0; t < taps; ++t) {
      const int64_t j = j_hi - t;
      if (j < 0) break;
      if (j < n_in) acc = fmaf(__ldg(h + t), load_mono(x, j, channels), acc);
    }
    out[o0 + n] = acc;
  }
}

}  // namespace svs

extern "C" int svs_resample_poly(const void* pcm, int pcm_is_int16, int channels, const int64_t* in_off,
                                 const int64_t* out_off, int n_songs, int64_t max_out, int up, int down,
                                 int64_t pre_pad, int64_t pre_remove, const float* h_poly, int taps, float* out,
                                 void* stream) {
  using namespace svs;
  SVS_REQUIRE(pcm && in_off && out_off && h_poly && out, "svs_resample_poly: null pointer");
  SVS_REQUIRE(n_songs > 0 && n_songs <= 65535 && channels >= 1 && channels <= 8, "svs_resample_poly: bad sizes");
  SVS_REQUIRE(up >= 1 && down >= 1 && taps >= 1 && max_out >= 0, "svs_resample_poly: bad rates");
  if (max_out == 0) return SVS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t blocks = (max_out + 255) / 256;
  const int64_t cap = (148 * 16 + n_songs - 1) / n_songs;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid(static_cast<unsigned>(blocks), n_songs);
  if (pcm_is_int16)
    resample_poly_kernel<int16_t><<<grid, 256, 0, st>>>(static_cast<const int16_t*>(pcm), channels, in_off, out_off, up,
                                                        down, pre_pad, pre_remove, h_poly, taps, out);
  else
    resample_poly_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(pcm), channels, in_off, out_off, up, down,
                                                      pre_pad, pre_remove, h_poly, taps, out);
  SVS_CHECK_LAUNCH("resample_poly_kernel");
  return SVS_OK;
}
