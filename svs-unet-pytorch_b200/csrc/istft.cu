// K2: batched inverse STFT with deterministic gather-form overlap-add for sm_100a.
//
// Replaces `librosa.istft(mag * phase, win_length=1024, hop_length=768)` (reference data.py:159)
// and the peak search of data.py:162: complex recombine fused into the load, inverse real FFT via
// the 512-point complex transform of fft512.cuh, Hann window, then each OUTPUT sample gathers the
// (at most two) frames covering it and divides by the window-sum-of-squares envelope.  There are
// no atomics on the waveform, so the result is bit-reproducible.
//
// With padded position P = p + 512:  frame t covers P in [768 t, 768 t + 1024).  Segment t =
// [768 t, 768 t + 768) receives frame t (offset r = P - 768 t) and, for r < 256, frame t-1
// (offset r + 768).  A 64-thread group owns a run of consecutive segments and transforms one extra
// frame (the one before the run, only for its tail), so there is no dependency between groups or CTAs.
#include "svs_common.cuh"
#include "fft512.cuh"

namespace svs {

constexpr int kIstftThreads = 256;
constexpr int kIstftFrFloats = 1024 + 128;                      // windowed frame, 4 floats of padding per 32 (z_addr)
constexpr int kIstftGroupFloats = 2 * kFftScratchFloats + 2 * kIstftFrFloats;   // exchange buffer X + two windowed frames
                                                                // (ping-pong); the current one doubles as exchange buffer Y
constexpr size_t kIstftSmemBytes = sizeof(float) * 4 * kIstftGroupFloats + sizeof(float2) * kFftTwiddleFloat2;

// The batch's frames are numbered globally (songs back to back, as frame_off lays them out) and split EVENLY over
// the 64-thread groups of a one-wave grid (3 CTAs per SM x 4 groups): group g walks the run [g R, (g+1) R) of
// consecutive frames, crossing song borders where they fall.  Frame t-1's last 256 windowed samples (its "tail") stay
// in shared memory and are added to the first 256 samples of frame t when hop segment t is emitted; the frame before
// a run is transformed only for its tail (one redundant transform per run, 0.6 % for the 150-song corpus), so groups
// never depend on each other and every output sample is written once.  (Round 1-2a: one CTA per 128 segments of one
// song, runs of 32 -> 3 % redundant transforms and 2,250 equal CTAs over 444 resident slots, i.e. SMs with 15 or 16 of
// them and a ragged last round.)
constexpr int kIstftMinRun = 8;

__global__ void __launch_bounds__(kIstftThreads, 3)
istft_ola_kernel(const float* __restrict__ mag, const float2* __restrict__ phase,
                 const int64_t* __restrict__ frame_off, const int64_t* __restrict__ wave_off, int n_songs,
                 float* __restrict__ wave, float* __restrict__ song_peak,
                 const float2* __restrict__ tw1024, const float* __restrict__ hann,
                 const float* __restrict__ env_both, const float* __restrict__ env_single) {
  extern __shared__ float smem[];
  const int group = threadIdx.x >> 6;
  const int j = threadIdx.x & 63;
  float2* tw_table = reinterpret_cast<float2*>(smem + 4 * kIstftGroupFloats);
  const FftTwiddles tw = build_fft_twiddles(tw_table, tw1024, threadIdx.x, kIstftThreads, j);
  __syncthreads();
  const int64_t first = frame_off[0], total = frame_off[n_songs];   // frames [first, total) of the mag / phase arrays
  const int64_t n_groups = static_cast<int64_t>(gridDim.x) * 4;
  int64_t run = (total - first + n_groups - 1) / n_groups;
  if (run < kIstftMinRun) run = kIstftMinRun;
  const int64_t g_begin = first + (static_cast<int64_t>(blockIdx.x) * 4 + group) * run;
  if (g_begin >= total) return;                                 // whole group leaves (barriers are per group)
  int remaining = static_cast<int>((g_begin + run <= total ? g_begin + run : total) - g_begin);   // segments to emit
  // song of the first frame: last s with frame_off[s] <= g_begin (songs without frames are skipped by the search)
  int song;
  {
    int lo = 0, hi = n_songs;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (frame_off[mid] <= g_begin) lo = mid; else hi = mid;
    }
    song = lo;
  }
  int64_t f0 = frame_off[song];
  int n_frames = static_cast<int>(frame_off[song + 1] - f0);
  float* scratch = smem + group * kIstftGroupFloats;
  float* xre = scratch;
  float* xim = scratch + kFftScratchFloats;
  // windowed frames ping-pong between two buffers: frame t-1's samples 768..1023 (its "tail") are read in place when
  // segment t is emitted -- no tail copy, and no barrier at the end of a frame (the buffer written next is the one
  // that was read two barriers-full frames ago)
  // Sample s lives at z_addr(s) = s + 4 (s >> 5): the producer's 8-byte stores (sample pairs 2n, n = jj + 64 d, i.e.
  // a stride of 16 floats across lanes) and the emitter's 4-byte loads of 32 consecutive samples are then both
  // bank-conflict free; the dense layout cost a 4-way conflict on every store (a third of all wavefronts: ncu).
  float* const fr_buf = scratch + 2 * kFftScratchFloats;        // [2][kIstftFrFloats]
  const int bar = 1 + group;

  float2 twp[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) twp[q] = __ldg(&tw1024[j + 64 * q]);
  // synthesis window x 1/512 for the samples this thread produces (n = jj + 64 d -> samples 2n, 2n + 1), in registers
  const int jj = (j >> 3) + 8 * (j & 7);
  // Samples 256..767 of a frame are covered by that frame alone, so their envelope reciprocal 1 / w^2 is folded into
  // the window factor here (one rounding apart from (x w) / w^2); samples < 256 and >= 768 overlap a neighbour and
  // keep the plain window until the two frames have been added.
  float2 wsc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    const int s0 = 2 * (jj + 64 * d);
    const float2 wv = __ldg(reinterpret_cast<const float2*>(hann) + jj + 64 * d);
    const bool single = s0 >= SVS_N_FFT - SVS_HOP && s0 < SVS_HOP;          // s0 even: s0 + 1 is in the same region
    const float e0 = single ? __ldg(&env_single[s0]) : 1.0f, e1 = single ? __ldg(&env_single[s0 + 1]) : 1.0f;
    wsc[d] = make_float2(wv.x * (1.0f / 512.0f) * e0, -wv.y * (1.0f / 512.0f) * e1);
  }
  // 1 / (w^2[r + 768] + w^2[r]) for the four samples r = 4 j .. 4 j + 3 < 256 this thread emits in the overlap region
  const float4 envb = __ldg(reinterpret_cast<const float4*>(env_both) + j);

  int64_t w0 = wave_off[song];
  int out_len = SVS_HOP * (n_frames - 1);                       // librosa: hop * (T - 1) after trimming
  float peak = 0.0f;
  auto flush_peak = [&]() {
    if (song_peak != nullptr) {
      const float m = warp_max(peak);
      if ((threadIdx.x & 31) == 0) atomic_max_nonneg(&song_peak[song], m);
    }
    peak = 0.0f;
  };

  // first iteration: the frame before the run, for its tail only -- or, at the start of a song, an empty tail
  int t = static_cast<int>(g_begin - f0);
  bool emit = t == 0;
  int par = 0;
  if (t == 0) {
    for (int i = j; i < 256; i += 64) fr_buf[kIstftFrFloats + z_addr(SVS_HOP + i)] = 0.0f;   // `prev` of parity 0
  } else {
    --t;
  }
  for (;;) {
    float* const fr = fr_buf + par * kIstftFrFloats;
    const float* const prev = fr_buf + (par ^ 1) * kIstftFrFloats;   // frame t-1: its samples 768..1023 are the tail
    const float* __restrict__ mrow = mag + (f0 + t) * SVS_N_BINS;
    const float2* __restrict__ prow = phase + (f0 + t) * SVS_N_BINS;
    if (f0 + t + 1 < total && j < 17 + 33) {                   // pull the next frame's 6,156 bytes (49 lines) into L2
      const char* nxt = j < 17 ? reinterpret_cast<const char*>(mrow + SVS_N_BINS) + 128 * j
                               : reinterpret_cast<const char*>(prow + SVS_N_BINS) + 128 * (j - 17);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
    }
    // Z[k] = E[k] + i O[k],  E = (X[k] + conj X[512-k])/2,  O = (X[k] - conj X[512-k])/2 * conj(W^k);
    // the inverse transform is conj(FFT(conj Z)), so conj(Z) is what goes into the exchange buffer.
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      if (q == 4 && j != 0) break;
      const int k = j + 64 * q;
      const int kk = 512 - k;
      const float mk = __ldg(&mrow[k]), mkk = __ldg(&mrow[kk]);
      const float2 pk = __ldg(&prow[k]), pkk = __ldg(&prow[kk]);
      float2 xk = make_float2(mk * pk.x, mk * pk.y);         // data.py:159  mag * phase
      float2 xkk = make_float2(mkk * pkk.x, mkk * pkk.y);
      if (k == 0) { xk.y = 0.0f; xkk.y = 0.0f; }             // c2r ignores Im of DC and Nyquist
      const float2 e = make_float2(0.5f * (xk.x + xkk.x), 0.5f * (xk.y - xkk.y));
      const float2 dd = make_float2(0.5f * (xk.x - xkk.x), 0.5f * (xk.y + xkk.y));
      const float2 w = (q < 4) ? twp[q & 3] : make_float2(0.0f, -1.0f);
      const float2 o = cmul(dd, cconj(w));
      // Z[k] = E + iO -> conj: (E.x - O.y, -(E.y + O.x)) ; Z[512-k] = conj(E) + i conj(O)
      const int ak = z_addr(k & 511);
      xre[ak] = e.x - o.y; xim[ak] = -(e.y + o.x);
      if (kk < 512 && kk != k) {
        const int akk = z_addr(kk);
        xre[akk] = e.x + o.y; xim[akk] = -(o.x - e.y);
      }
    }
    group_bar(bar);
    float2 v[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const int a = z_addr(j + 64 * n1);
      v[n1] = make_float2(xre[a], xim[a]);
    }
    // The FFT's first exchange goes through this frame's (still empty) `fr` buffer and its second one through X, so
    // neither the read of conj(Z) above nor the windowed stores below need a barrier of their own: pass A writes a
    // buffer nobody reads until its own barrier, and `fr` was last read (pass B) before the FFT's second barrier.
    // Four group barriers per frame instead of six.
    fft512_group(v, tw, fr, scratch, j, bar);
    // v[d] = FFT(conj Z)[n], n = jj + 64 d ;  z[n] = conj(v)/512 ;  x[2n] = Re, x[2n+1] = Im
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const int n = jj + 64 * d;
      *reinterpret_cast<float2*>(&fr[z_addr(2 * n)]) = make_float2(v[d].x * wsc[d].x, v[d].y * wsc[d].y);
    }
    group_bar(bar);
    // ---- emit hop segment t (gather form: frame t-1's tail + frame t), then keep frame t's tail ----
    // A thread emits FOUR consecutive samples per step (r = 4 j + 256 c): 16-byte shared-memory loads (four samples
    // never straddle the 4-float padding every 32) and 16-byte global stores, 3 + 1 loads and 3 stores per frame
    // instead of 12 + 4 and 12.  Segment borders (p = 0, out_len) are multiples of 4, so a quad is all in or all out.
    if (emit) {
#pragma unroll
      for (int c = 0; c < SVS_HOP / 256; ++c) {
        const int r = 4 * j + 256 * c;
        const int p = t * SVS_HOP + r - SVS_N_FFT / 2;        // first of the four output samples
        if (p >= 0 && p < out_len) {
          float4 v = *reinterpret_cast<const float4*>(&fr[z_addr(r)]);
          if (c == 0) {
            // r < 256: frame t-1's tail is added first, then the two-frame envelope (t = 0 never gets here: p < 0);
            // r >= 256: already divided by its envelope when it was windowed
            const float4 tl = *reinterpret_cast<const float4*>(&prev[z_addr(SVS_HOP + r)]);
            v.x = (tl.x + v.x) * envb.x; v.y = (tl.y + v.y) * envb.y;
            v.z = (tl.z + v.z) * envb.z; v.w = (tl.w + v.w) * envb.w;
          }
          *reinterpret_cast<float4*>(&wave[w0 + p]) = v;
          peak = fmaxf(peak, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
      }
      if (--remaining == 0) break;
    }
    emit = true;
    par ^= 1;
    if (++t >= n_frames) {                                   // the run continues in the next song
      flush_peak();
      do {
        ++song;
        f0 = frame_off[song];
        n_frames = static_cast<int>(frame_off[song + 1] - f0);
      } while (n_frames <= 0);                               // remaining > 0 guarantees a frame further on
      w0 = wave_off[song];
      out_len = SVS_HOP * (n_frames - 1);
      t = 0;
      // empty tail for the new song's first segment: `prev` of the next iteration is the buffer just emitted from;
      // its tail region is read by nobody any more, and the next frame's barriers order these stores before its emit
      for (int i = j; i < 256; i += 64) fr_buf[(par ^ 1) * kIstftFrFloats + z_addr(SVS_HOP + i)] = 0.0f;
    }
  }
  flush_peak();
}

// grid (blocks per song, n_songs): no per-element song lookup, 16-byte accesses (song offsets are multiples of the
// 768-sample hop, so a song starts 16-byte aligned whenever the buffer does)
__global__ void __launch_bounds__(256)
wave_peak_normalize_kernel(float* __restrict__ wave, const int64_t* __restrict__ wave_off,
                           const float* __restrict__ song_peak, float target) {
  const int s = blockIdx.y;
  const float pk = song_peak[s];
  if (!(pk > 0.0f)) return;                                  // reference data.py:163: silent songs stay as they are
  const int64_t a = wave_off[s], n = wave_off[s + 1] - a;
  float* w = wave + a;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    float4* w4 = reinterpret_cast<float4*>(w);
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      float4 v = w4[i];
      v.x = v.x / pk * target; v.y = v.y / pk * target; v.z = v.z / pk * target; v.w = v.w / pk * target;   // data.py:164
      w4[i] = v;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) w[i] = w[i] / pk * target;
  } else {
    for (int64_t i = tid; i < n; i += stride) w[i] = w[i] / pk * target;
  }
}

// The PCM_16 writer of reference data.py:166 (soundfile.write -> libsndfile f2les_array: lrintf(x * 0x7FFF)) fused
// with the 0.9 / peak normalisation of data.py:163-164: one pass, 4 bytes read + 2 bytes written per sample, and the
// download is half the size of the float waveform.
__global__ void __launch_bounds__(256)
wave_peak_normalize_pcm16_kernel(const float* __restrict__ wave, const int64_t* __restrict__ wave_off,
                                 const float* __restrict__ song_peak, float target, int16_t* __restrict__ out) {
  const int s = blockIdx.y;
  const float pk = song_peak[s];
  const bool scale = pk > 0.0f;                              // data.py:163: silent songs stay as they are
  const int64_t a = wave_off[s], n = wave_off[s + 1] - a;
  const float* w = wave + a;
  int16_t* o = out + a;
  auto q = [&](float v) {
    if (scale) v = v / pk * target;                          // data.py:164, same expression as the float pass
    const float r = fminf(fmaxf(v * 32767.0f, -32768.0f), 32767.0f);
    return static_cast<int16_t>(__float2int_rn(r));
  };
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 7) == 0) {
    const float4* w4 = reinterpret_cast<const float4*>(w);
    short4* o4 = reinterpret_cast<short4*>(o);
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 v = __ldcs(&w4[i]);
      o4[i] = make_short4(q(v.x), q(v.y), q(v.z), q(v.w));
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) o[i] = q(w[i]);
  } else {
    for (int64_t i = tid; i < n; i += stride) o[i] = q(w[i]);
  }
}

}  // namespace svs

extern "C" int svs_istft_ola(const float* mag, const float* phase, const int64_t* frame_off,
                             const int64_t* wave_off, int n_songs, int64_t max_frames, float* wave,
                             float* song_peak, void* stream) {
  using namespace svs;
  SVS_REQUIRE(mag && phase && frame_off && wave_off && wave, "svs_istft_ola: null pointer");
  SVS_REQUIRE(n_songs > 0 && n_songs <= 65535, "svs_istft_ola: n_songs must be in [1, 65535]");
  SVS_REQUIRE(max_frames > 0, "svs_istft_ola: max_frames must be positive");
  SVS_REQUIRE((reinterpret_cast<uintptr_t>(wave) & 15) == 0, "svs_istft_ola: wave must be 16-byte aligned");
  SpectralTables tabs;
  int rc = get_spectral_tables(&tabs);
  if (rc != SVS_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SVS_CUDA_TRY(cudaFuncSetAttribute(istft_ola_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(kIstftSmemBytes)));
  if (song_peak) SVS_CUDA_TRY(cudaMemsetAsync(song_peak, 0, sizeof(float) * n_songs, st));
  // one wave: 3 CTAs per SM (or fewer when the batch is small: a group emits at least kIstftMinRun segments)
  const int64_t frames_bound = max_frames * n_songs;
  int64_t ctas = static_cast<int64_t>(num_sms()) * 3;
  const int64_t need = (frames_bound + 4 * kIstftMinRun - 1) / (4 * kIstftMinRun);
  if (ctas > need) ctas = need;
  if (ctas < 1) ctas = 1;
  istft_ola_kernel<<<static_cast<unsigned>(ctas), kIstftThreads, kIstftSmemBytes, st>>>(
      mag, reinterpret_cast<const float2*>(phase), frame_off, wave_off, n_songs, wave, song_peak, tabs.tw1024,
      tabs.hann, tabs.env_both, tabs.env_single);
  SVS_CHECK_LAUNCH("istft_ola_kernel");
  return SVS_OK;
}

extern "C" int svs_wave_peak_normalize(float* wave, const int64_t* wave_off, const float* song_peak,
                                       int n_songs, int64_t total_samples, float target, void* stream) {
  using namespace svs;
  SVS_REQUIRE(wave && wave_off && song_peak, "svs_wave_peak_normalize: null pointer");
  SVS_REQUIRE(n_songs > 0 && n_songs <= 65535 && total_samples >= 0, "svs_wave_peak_normalize: bad sizes");
  if (total_samples == 0) return SVS_OK;
  // enough CTAs to fill the GPU for a single song, fewer per song when there are many
  int per_song = (148 * 8 + n_songs - 1) / n_songs;
  if (per_song < 4) per_song = 4;
  wave_peak_normalize_kernel<<<dim3(per_song, n_songs), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      wave, wave_off, song_peak, target);
  SVS_CHECK_LAUNCH("wave_peak_normalize_kernel");
  return SVS_OK;
}

extern "C" int svs_wave_peak_normalize_pcm16(const float* wave, const int64_t* wave_off, const float* song_peak,
                                             int n_songs, int64_t total_samples, float target, int16_t* pcm_out,
                                             void* stream) {
  using namespace svs;
  SVS_REQUIRE(wave && wave_off && song_peak && pcm_out, "svs_wave_peak_normalize_pcm16: null pointer");
  SVS_REQUIRE(n_songs > 0 && n_songs <= 65535 && total_samples >= 0, "svs_wave_peak_normalize_pcm16: bad sizes");
  if (total_samples == 0) return SVS_OK;
  int per_song = (148 * 8 + n_songs - 1) / n_songs;
  if (per_song < 4) per_song = 4;
  wave_peak_normalize_pcm16_kernel<<<dim3(per_song, n_songs), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      wave, wave_off, song_peak, target, pcm_out);
  SVS_CHECK_LAUNCH("wave_peak_normalize_pcm16_kernel");
  return SVS_OK;
}
