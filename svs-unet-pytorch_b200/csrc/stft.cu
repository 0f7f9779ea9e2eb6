// K1: batched magnitude/phase STFT for sm_100a.
//
// Replaces librosa.stft(n_fft=1024, hop=768, center=True, pad_mode="constant") + librosa.magphase
// (reference data.py:79-81, 100-102) and the per-song max of data.py:84.
//
// HBM-bound (9,228 algorithmic bytes/frame, ~35 kFLOP/frame): the Hann window is fused into the
// load (centre zero padding = predicated loads, no padded copy), the 1024-point real FFT is a
// 512-point complex FFT in registers + conflict-free shared-memory exchanges (fft512.cuh), and
// magnitude / unit phasor / per-song max are fused into the store.  Output rows are frames
// ([T][513]), which is byte-for-byte the Fortran-ordered (513, T) array librosa returns.
//
// Grid: (frame chunks, songs); CTA = 256 threads = 4 groups of 64, one frame per group at a time.
#include "svs_common.cuh"
#include "fft512.cuh"

namespace svs {

constexpr int kStftThreads = 256;
constexpr int kStftFramesPerCta = 64;     // 16 frames per 64-thread group: the per-CTA prologue (26 twiddle / window loads per
                                          // thread) is worth about one frame; 32 per CTA cost 3 points of HBM utilisation

// |x| and x/|x| (1+0j where |x| == 0, librosa.magphase) with one rsqrt instead of a sqrt and two divisions
__device__ __forceinline__ void mag_phase(float2 x, float& m, float2& ph) {
  const float m2 = x.x * x.x + x.y * x.y;
  if (m2 > 1e-37f) {
    const float inv = rsqrtf(m2);
    m = m2 * inv;
    ph = make_float2(x.x * inv, x.y * inv);
  } else {                                   // zero / denormal energy: exact path (reference data.py:80 semantics)
    m = sqrtf(m2);
    const float z = (m == 0.0f) ? 1.0f : 0.0f;
    ph = make_float2(x.x / (m + z) + z, x.y / (m + z));
  }
}

// InT = float (librosa.load's output) or int16_t (the PCM_16 samples of the .wav itself, reference data.py:78:
// libsndfile / librosa.load scale them by 1/32768 — fused into the load, so the upload is 2 bytes per sample)
template <bool kComplexOut, typename InT>
__global__ void __launch_bounds__(kStftThreads, 3)
stft_mag_phase_kernel(const InT* __restrict__ audio, const int64_t* __restrict__ sample_off,
                      const int64_t* __restrict__ frame_off, float* __restrict__ mag,
                      float2* __restrict__ phase, float* __restrict__ song_max,
                      const float2* __restrict__ tw1024, const float* __restrict__ hann) {
  __shared__ float scratch_all[4 * kFftGroupFloats];
  const int song = blockIdx.y;
  const int64_t f0 = frame_off[song];
  const int n_frames = static_cast<int>(frame_off[song + 1] - f0);
  const int t_begin = blockIdx.x * kStftFramesPerCta;
  if (t_begin >= n_frames) return;
  const int t_end = min(n_frames, t_begin + kStftFramesPerCta);
  const int64_t s0 = sample_off[song];
  const int len = static_cast<int>(sample_off[song + 1] - s0);
  const InT* __restrict__ y = audio + s0;
  const bool vec2 = (reinterpret_cast<uintptr_t>(y) & (2 * sizeof(InT) - 1)) == 0;   // frame starts are even sample offsets

  const int group = threadIdx.x >> 6;
  const int j = threadIdx.x & 63;
  float* scratch = scratch_all + group * kFftGroupFloats;
  const int bar = 1 + group;

  // per-thread constants in registers: window taps, FFT twiddles, split-step twiddles (the kernel's limiter
  // is shared-memory bandwidth, so tables in shared memory cost more than the occupancy they buy: measured)
  __shared__ float2 tw_table[kFftTwiddleFloat2];
  const FftTwiddles tw = build_fft_twiddles(tw_table, tw1024, threadIdx.x, kStftThreads, j);
  __syncthreads();
  float2 win[8];
#pragma unroll
  for (int n1 = 0; n1 < 8; ++n1) win[n1] = __ldg(reinterpret_cast<const float2*>(hann) + j + 64 * n1);
  float2 twp[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) twp[q] = __ldg(&tw1024[j + 64 * q]);

  float* xre = scratch;
  float* xim = scratch + kFftScratchFloats;
  float run_max = 0.0f;

  // software pipeline: the raw samples of this group's NEXT frame are requested before the current frame's FFT,
  // so their DRAM latency overlaps the three radix passes (16 warps / SM cannot hide it by occupancy alone)
  auto load_frame = [&](int t, float2 (&raw)[8]) {
    const int base = t * SVS_HOP - SVS_N_FFT / 2;          // frame t covers samples [768 t - 512, 768 t + 512)
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const int i0 = base + 2 * (j + 64 * n1);
      if constexpr (sizeof(InT) == 4) {
        if (vec2 && i0 >= 0 && i0 + 1 < len) {
          raw[n1] = __ldg(reinterpret_cast<const float2*>(y + i0));
        } else {
          raw[n1].x = (i0 >= 0 && i0 < len) ? __ldg(&y[i0]) : 0.0f;
          raw[n1].y = (i0 + 1 >= 0 && i0 + 1 < len) ? __ldg(&y[i0 + 1]) : 0.0f;
        }
      } else {
        constexpr float kScale = 1.0f / 32768.0f;
        if (vec2 && i0 >= 0 && i0 + 1 < len) {
          const short2 q = __ldg(reinterpret_cast<const short2*>(y + i0));
          raw[n1] = make_float2(static_cast<float>(q.x) * kScale, static_cast<float>(q.y) * kScale);
        } else {
          raw[n1].x = (i0 >= 0 && i0 < len) ? static_cast<float>(__ldg(&y[i0])) * kScale : 0.0f;
          raw[n1].y = (i0 + 1 >= 0 && i0 + 1 < len) ? static_cast<float>(__ldg(&y[i0 + 1])) * kScale : 0.0f;
        }
      }
    }
  };
  float2 nxt[8];
  if (t_begin + group < t_end) load_frame(t_begin + group, nxt);
  for (int t = t_begin + group; t < t_end; t += 4) {
    float2 v[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) v[n1] = make_float2(nxt[n1].x * win[n1].x, nxt[n1].y * win[n1].y);
    if (t + 4 < t_end) load_frame(t + 4, nxt);
    fft512_group(v, tw, scratch, scratch + 2 * kFftScratchFloats, j, bar);
    // ---- exchange 3: Z[k] in padded linear order ----
    const int jj = (j >> 3) + 8 * (j & 7);
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const int a = z_addr(jj + 64 * d);
      xre[a] = v[d].x; xim[a] = v[d].y;
    }
    group_bar(bar);
    // ---- split step + magnitude / phase store ----
    float* __restrict__ mrow = mag + (f0 + t) * SVS_N_BINS;
    float2* __restrict__ prow = phase ? phase + (f0 + t) * SVS_N_BINS : nullptr;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      if (q == 4 && j != 0) break;
      const int k = j + 64 * q;
      const int kk = 512 - k;
      const int ak = z_addr(k & 511), akk = z_addr(kk & 511);
      const float2 zk = make_float2(xre[ak], xim[ak]);
      const float2 zkk = make_float2(xre[akk], xim[akk]);
      // E = (Zk + conj Zkk)/2 ; O = -(i/2)(Zk - conj Zkk) ; X[k] = E + W^k O ; X[512-k] = conj(E - W^k O)
      const float2 e = make_float2(0.5f * (zk.x + zkk.x), 0.5f * (zk.y - zkk.y));
      const float2 o = make_float2(0.5f * (zk.y + zkk.y), -0.5f * (zk.x - zkk.x));
      const float2 w = (q < 4) ? twp[q & 3] : make_float2(0.0f, -1.0f);   // W_1024^256 = -i
      const float2 tt = cmul(w, o);
      const float2 xk = cadd(e, tt);
      const float2 xkk = cconj(csub(e, tt));
      if constexpr (kComplexOut) {                             // raw spectrum (librosa.stft drop-in)
        prow[k] = xk;
        if (kk != k) prow[kk] = xkk;
      } else {
        {
          float m; float2 ph;
          mag_phase(xk, m, ph);
          mrow[k] = m;
          run_max = fmaxf(run_max, m);
          if (prow) prow[k] = ph;
        }
        if (kk != k) {
          float m; float2 ph;
          mag_phase(xkk, m, ph);
          mrow[kk] = m;
          run_max = fmaxf(run_max, m);
          if (prow) prow[kk] = ph;
        }
      }
    }
    group_bar(bar);   // buffer X is rewritten by the next frame's pass A
  }
  if (song_max != nullptr) {
    run_max = warp_max(run_max);
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg(&song_max[song], run_max);
  }
}

// librosa.magphase on an arbitrary complex64 array (reference data.py:80,101)
__global__ void magphase_kernel(const float2* __restrict__ d, int64_t n, float* __restrict__ mag,
                                float2* __restrict__ phase) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float2 x = d[i];
    const float m = sqrtf(x.x * x.x + x.y * x.y);
    const float z = (m == 0.0f) ? 1.0f : 0.0f;
    mag[i] = m;
    phase[i] = make_float2(x.x / (m + z) + z, x.y / (m + z));
  }
}

__global__ void spec_normalize_kernel(float* __restrict__ mag, const int64_t* __restrict__ frame_off,
                                      const float* __restrict__ norm, int n_songs, int64_t total_frames) {
  // one CTA row per frame keeps the song lookup out of the element loop
  for (int64_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
    int lo = 0, hi = n_songs - 1;
    while (lo < hi) {                       // last song with frame_off[s] <= f
      const int mid = (lo + hi + 1) >> 1;
      if (frame_off[mid] <= f) lo = mid; else hi = mid - 1;
    }
    float nrm = norm[lo];
    if (nrm == 0.0f) nrm = 1.0f;            // reference data.py:85
    float* row = mag + f * SVS_N_BINS;
    for (int k = threadIdx.x; k < SVS_N_BINS; k += blockDim.x) row[k] = row[k] / nrm;
  }
}

}  // namespace svs

namespace svs {
template <typename InT>
static int launch_stft(const InT* audio, const int64_t* sample_off, const int64_t* frame_off, int n_songs,
                       int64_t max_frames, float* mag, float* phase, float* song_max, void* stream, const char* what) {
  if (!(audio && sample_off && frame_off && mag)) return fail(SVS_ERR_INVALID_ARG, std::string(what) + ": null pointer");
  if (!(n_songs > 0 && n_songs <= 65535)) return fail(SVS_ERR_INVALID_ARG, std::string(what) + ": n_songs must be in [1, 65535]");
  if (!(max_frames > 0)) return fail(SVS_ERR_INVALID_ARG, std::string(what) + ": max_frames must be positive");
  SpectralTables tabs;
  int rc = get_spectral_tables(&tabs);
  if (rc != SVS_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (song_max) SVS_CUDA_TRY(cudaMemsetAsync(song_max, 0, sizeof(float) * n_songs, st));
  dim3 grid(static_cast<unsigned>((max_frames + kStftFramesPerCta - 1) / kStftFramesPerCta), n_songs);
  stft_mag_phase_kernel<false, InT><<<grid, kStftThreads, 0, st>>>(audio, sample_off, frame_off, mag,
                                                                   reinterpret_cast<float2*>(phase), song_max,
                                                                   tabs.tw1024, tabs.hann);
  SVS_CHECK_LAUNCH("stft_mag_phase_kernel");
  return SVS_OK;
}
}  // namespace svs

extern "C" int svs_stft_mag_phase(const float* audio, const int64_t* sample_off, const int64_t* frame_off,
                                  int n_songs, int64_t max_frames, float* mag, float* phase,
                                  float* song_max, void* stream) {
  return svs::launch_stft(audio, sample_off, frame_off, n_songs, max_frames, mag, phase, song_max, stream,
                          "svs_stft_mag_phase");
}

extern "C" int svs_stft_mag_phase_pcm16(const int16_t* audio, const int64_t* sample_off, const int64_t* frame_off,
                                        int n_songs, int64_t max_frames, float* mag, float* phase,
                                        float* song_max, void* stream) {
  return svs::launch_stft(audio, sample_off, frame_off, n_songs, max_frames, mag, phase, song_max, stream,
                          "svs_stft_mag_phase_pcm16");
}

extern "C" int svs_stft_complex(const float* audio, const int64_t* sample_off, const int64_t* frame_off,
                                int n_songs, int64_t max_frames, float* spec, void* stream) {
  using namespace svs;
  SVS_REQUIRE(audio && sample_off && frame_off && spec, "svs_stft_complex: null pointer");
  SVS_REQUIRE(n_songs > 0 && n_songs <= 65535, "svs_stft_complex: n_songs must be in [1, 65535]");
  SVS_REQUIRE(max_frames > 0, "svs_stft_complex: max_frames must be positive");
  SpectralTables tabs;
  int rc = get_spectral_tables(&tabs);
  if (rc != SVS_OK) return rc;
  dim3 grid(static_cast<unsigned>((max_frames + kStftFramesPerCta - 1) / kStftFramesPerCta), n_songs);
  stft_mag_phase_kernel<true, float><<<grid, kStftThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      audio, sample_off, frame_off, nullptr, reinterpret_cast<float2*>(spec), nullptr, tabs.tw1024, tabs.hann);
  SVS_CHECK_LAUNCH("stft_complex_kernel");
  return SVS_OK;
}

extern "C" int svs_magphase(const float* spec, int64_t n, float* mag, float* phase, void* stream) {
  using namespace svs;
  SVS_REQUIRE(spec && mag && phase && n >= 0, "svs_magphase: bad arguments");
  if (n == 0) return SVS_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  magphase_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(spec), n, mag, reinterpret_cast<float2*>(phase));
  SVS_CHECK_LAUNCH("magphase_kernel");
  return SVS_OK;
}

extern "C" int svs_spec_normalize(float* mag, const int64_t* frame_off, const float* norm, int n_songs,
                                  int64_t total_frames, void* stream) {
  using namespace svs;
  SVS_REQUIRE(mag && frame_off && norm, "svs_spec_normalize: null pointer");
  SVS_REQUIRE(n_songs > 0 && total_frames >= 0, "svs_spec_normalize: bad sizes");
  if (total_frames == 0) return SVS_OK;
  const int64_t blocks = total_frames < 148 * 16 ? total_frames : 148 * 16;
  spec_normalize_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      mag, frame_off, norm, n_songs, total_frames);
  SVS_CHECK_LAUNCH("spec_normalize_kernel");
  return SVS_OK;
}
