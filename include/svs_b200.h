/*
 * svs_b200.h — C ABI of libsvs_b200.so, the B200 (sm_100a) implementation of the SVS-UNet
 * separation hot path:  magnitude STFT -> UNet soft mask -> mask x mixture -> iSTFT overlap-add,
 * plus the L1 training step through the same kernels.
 *
 * The reference (zouyuoz/SVS-UNet-PyTorch) has no FFI of its own: its hot path is reached through
 * librosa (data.py) and torch.nn (model.py).  Each entry point below names the reference call
 * site(s) it replaces.  The binding a maintainer adds on the reference side is a ctypes stub —
 * see INTEGRATION.md.
 *
 * Conventions
 *   - every call returns 0 on success, a negative svs_status otherwise; the message is available
 *     from svs_last_error() (thread local).  No C++ exception crosses this boundary.
 *   - all data pointers are DEVICE pointers owned by the caller; the library never allocates or
 *     frees caller-visible memory (a plan owns only its private repacked weights).
 *   - all work is enqueued on the caller's cudaStream_t (passed as void*); calls are asynchronous,
 *     never synchronise the device and are CUDA-graph capturable.
 *   - there is no CPU path and no fallback: a device that is not compute capability 10.x, or a
 *     geometry other than n_fft 1024 / hop 768 / 512x128 patches, is an error.
 */
#ifndef SVS_B200_H
#define SVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVS_ABI_VERSION 3   /* 3: + svs_patch_stream_* (host-buffer patch batches) */

/* geometry of the path (reference config.py:47-51) */
#define SVS_N_FFT        1024
#define SVS_HOP          768
#define SVS_N_BINS       513      /* rows of a *_spec.npy */
#define SVS_PATCH_BINS   512      /* DC row dropped, reference inference.py:68 */
#define SVS_PATCH_FRAMES 128      /* INPUT_LEN */

typedef enum svs_status {
  SVS_OK = 0,
  SVS_ERR_INVALID_ARG = -1,
  SVS_ERR_CUDA = -2,
  SVS_ERR_UNSUPPORTED_ARCH = -3,
  SVS_ERR_WORKSPACE = -4,
  SVS_ERR_NOT_IMPLEMENTED = -5
} svs_status;

typedef enum svs_precision {
  SVS_PRECISION_FP32 = 0,   /* fp32 activations + fp32 FMA (exact-arithmetic mode)            */
  SVS_PRECISION_BF16 = 1,   /* bf16 NHWC activations, tcgen05 kind::f16 MMA, fp32 accumulate   */
  SVS_PRECISION_TF32 = 2    /* fp32 NHWC activations, tcgen05 kind::tf32 MMA, fp32 accumulate  */
} svs_precision;

enum {
  SVS_FLAG_APPLY_MASK = 1,  /* out = mix * mask (reference inference.py:107) instead of the mask   */
  SVS_FLAG_INVERT     = 2   /* mask <- 1 - mask  (reference inference.py:102, --vocal_solo 0)       */
};

/* ------------------------------------------------------------------ lifecycle / diagnostics */
int svs_version(void);                       /* == SVS_ABI_VERSION                              */
const char* svs_last_error(void);            /* thread-local message of the last failing call   */
int svs_device_check(int device);            /* SVS_OK iff `device` is compute capability 10.x  */

/* ------------------------------------------------------------------ decode side
 * The arithmetic of `librosa.load(path, sr=8192, mono=True)` (reference data.py:78,94) after the file bytes are
 * read: PCM -> float (int16 / 32768), channel mean, polyphase rational resampling with scipy.signal.resample_poly's
 * conventions (librosa's own soxr_hq filter is not reproducible: see csrc/resample.cu).
 *   pcm        interleaved frames of `channels` samples, int16 or float32; song s = frames in_off[s] .. in_off[s+1])
 *   out        float32 mono; song s = out[out_off[s] .. out_off[s+1]), out length = ceil(n_in * up / down)
 *   h_poly     device float32 [up][taps]: h_poly[p][t] = h[p + t * up] of the (gain `up`) low-pass h
 *   pre_pad, pre_remove   scipy's n_pre_pad / n_pre_remove for (len(h), up, down) */
int svs_resample_poly(const void* pcm, int pcm_is_int16, int channels, const int64_t* in_off,
                      const int64_t* out_off, int n_songs, int64_t max_out, int up, int down, int64_t pre_pad,
                      int64_t pre_remove, const float* h_poly, int taps, float* out, void* stream);

/* ------------------------------------------------------------------ spectral front end (S1-S3)
 * Replaces librosa.stft + librosa.magphase at reference data.py:79-81 and data.py:100-102
 * (n_fft 1024, hop 768, periodic Hann, center=True, pad_mode="constant"), batched over a ragged
 * set of songs.
 *   audio        concatenated mono float32 songs; song s = audio[sample_off[s] .. sample_off[s+1])
 *   sample_off   device int64 [n_songs+1]
 *   frame_off    device int64 [n_songs+1]; frame_off[s+1]-frame_off[s] must equal 1 + len_s/768
 *   max_frames   host: largest per-song frame count (grid sizing)
 *   mag          float32 [total_frames][513]  == the Fortran-ordered (513,T) array librosa returns
 *   phase        complex64 as float pairs [total_frames][513][2], unit phasors, 1+0j where mag==0;
 *                may be NULL
 *   song_max     float32 [n_songs] max magnitude per song (reference data.py:84); may be NULL
 */
int svs_stft_mag_phase(const float* audio, const int64_t* sample_off, const int64_t* frame_off,
                       int n_songs, int64_t max_frames, float* mag, float* phase, float* song_max,
                       void* stream);

/* The same transform straight from the PCM_16 samples of the .wav file (reference data.py:78: librosa.load of a
 * 16-bit file is int16 / 32768, applied here inside the load): `audio` is device int16, song s =
 * audio[sample_off[s] .. sample_off[s+1]).  Halves the host -> device bytes of the end-to-end path. */
int svs_stft_mag_phase_pcm16(const int16_t* audio, const int64_t* sample_off, const int64_t* frame_off,
                             int n_songs, int64_t max_frames, float* mag, float* phase, float* song_max,
                             void* stream);

/* librosa.stft alone (reference data.py:79,100): raw complex64 spectrum [total_frames][513][2]. */
int svs_stft_complex(const float* audio, const int64_t* sample_off, const int64_t* frame_off,
                     int n_songs, int64_t max_frames, float* spec, void* stream);

/* librosa.magphase alone (reference data.py:80,101) on n complex64 values. */
int svs_magphase(const float* spec, int64_t n, float* mag, float* phase, void* stream);

/* spec /= norm per song, norm==0 -> 1 (reference data.py:85,105).  `norm` is a device float32
 * [n_songs] array (for the vocal stem it is the MIXTURE's song_max). */
int svs_spec_normalize(float* mag, const int64_t* frame_off, const float* norm, int n_songs,
                       int64_t total_frames, void* stream);

/* ------------------------------------------------------------------ spectral back end (W1-W3)
 * Replaces `librosa.istft(mag * phase, win_length=1024, hop_length=768)` at reference data.py:159:
 * complex recombine, inverse real FFT, Hann window, deterministic gather-form overlap-add (no
 * atomics), division by the window-sum-of-squares envelope, trim n_fft/2 on both ends.
 *   wave         float32, 16-byte aligned; song s occupies wave[wave_off[s] .. wave_off[s] + 768*(T_s-1)), wave_off[s]
 *                a multiple of 4 (back-to-back songs: multiples of 768)
 *   song_peak    float32 [n_songs] max |y| per song (reference data.py:162); may be NULL
 */
int svs_istft_ola(const float* mag, const float* phase, const int64_t* frame_off,
                  const int64_t* wave_off, int n_songs, int64_t max_frames, float* wave,
                  float* song_peak, void* stream);

/* y <- y / peak * target where peak > 0 (reference data.py:163-164, target 0.9). */
int svs_wave_peak_normalize(float* wave, const int64_t* wave_off, const float* song_peak,
                            int n_songs, int64_t total_samples, float target, void* stream);

/* The normalisation above fused with the PCM_16 quantiser of `sf.write(path, y, sr)` at reference data.py:166
 * (libsndfile: lrintf(y * 0x7FFF)): pcm_out[i] = rint(wave[i] / peak * target * 32767), int16, same offsets as
 * `wave` (which is left untouched).  Halves the device -> host bytes of the end-to-end path. */
int svs_wave_peak_normalize_pcm16(const float* wave, const int64_t* wave_off, const float* song_peak,
                                  int n_songs, int64_t total_samples, float target, int16_t* pcm_out,
                                  void* stream);

/* ------------------------------------------------------------------ patch staging (I2, I4)
 * Replaces the per-patch segment / zero-pad / contiguous copy of reference inference.py:74-97 and the
 * crop / concatenate / DC re-insert of inference.py:110-127, for `n` patches at once.
 *   spec        frame-major song spectrogram(s) [frames][513] float32
 *   patch_off   device [n] element offsets of each patch's first element (frame * 513 + 1: DC bin skipped)
 *   in_frames   device [n] valid frames per patch (<= 128), or NULL for 128
 *   patches     dense [n][512][128] float32
 * gather: patches[p][f][t] = spec[patch_off[p] + t*513 + f] / norm[p] for t < in_frames[p], else 0; `norm` (device
 * [n], NULL = 1) folds the per-song normalisation of data.py:85,105 (0 -> 1) into the copy.
 * scatter: the inverse for t < in_frames[p]; dc_zero != 0 also writes 0 to the DC bin of those frames. */
int svs_patches_gather(const float* spec, const int64_t* patch_off, const int32_t* in_frames,
                       const float* norm, float* patches, int n, void* stream);
int svs_patches_scatter(const float* patches, const int64_t* patch_off, const int32_t* in_frames,
                        float* spec, int n, int dc_zero, void* stream);

/* ------------------------------------------------------------------ UNet mask (U1-D6, I2-I4)
 * Replaces UNet.forward (reference model.py:169-201) in eval mode followed by the mask
 * application of reference inference.py:102,107.
 */
typedef struct svs_conv_params {      /* one conv / deconv block, tensors in torch layout, fp32 */
  const float* weight;                /* Conv2d (Cout,Cin,5,5) / ConvTranspose2d (Cin,Cout,5,5) */
  const float* bias;                  /* (Cout)                                                  */
  const float* bn_weight;             /* (Cout) or NULL when the block has no BatchNorm (deconv6)*/
  const float* bn_bias;
  const float* bn_mean;               /* running_mean */
  const float* bn_var;                /* running_var  */
} svs_conv_params;

typedef struct svs_unet_plan svs_unet_plan;   /* opaque; immutable after creation */

/* layers[0..5] = conv1..conv6, layers[6..11] = deconv1..deconv6 (state_dict of model.py:47-109).
 * Folds eval-mode BatchNorm (eps 1e-5) into weights/bias and repacks them for the kernels. The
 * source tensors may be freed once the call's stream work has completed. */
int svs_unet_plan_create(const svs_conv_params layers[12], int precision, void* stream,
                         svs_unet_plan** plan_out);
int svs_unet_plan_destroy(svs_unet_plan* plan);
int svs_unet_plan_precision(const svs_unet_plan* plan);

size_t svs_unet_workspace_bytes(const svs_unet_plan* plan, int batch);

/* One batch of `batch` patches, each 512 bins x 128 frames.
 *   in / out      float32; element (patch b, bin f, frame t) lives at
 *                 base + patch_off[b] + f*stride_f + t*stride_t   (patch_off == NULL: b*stride_b)
 *   in_frames     device int32 [batch] number of valid frames per patch (the rest are read as zero
 *                 and not written: the zero padding / crop of reference inference.py:90-92,113-114);
 *                 NULL = all 128
 *   flags         SVS_FLAG_APPLY_MASK | SVS_FLAG_INVERT
 *   workspace     device scratch of at least svs_unet_workspace_bytes(plan, batch), 1024-B aligned
 */
typedef struct svs_patch_view {
  float* base;
  const int64_t* patch_off;   /* device [batch] element offsets, or NULL */
  int64_t stride_b;           /* used when patch_off == NULL */
  int64_t stride_f;
  int64_t stride_t;
} svs_patch_view;

int svs_unet_forward(const svs_unet_plan* plan, const svs_patch_view* in, const svs_patch_view* out,
                     const int32_t* in_frames, int batch, int flags, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Profiling hook: enqueue only layers [first_layer, last_layer] (0 = conv1 .. 11 = deconv6) on a
 * workspace that already holds the inputs of `first_layer` from an earlier full forward. */
int svs_unet_forward_layers(const svs_unet_plan* plan, const svs_patch_view* in, const svs_patch_view* out,
                            const int32_t* in_frames, int batch, int flags, void* workspace,
                            size_t workspace_bytes, int first_layer, int last_layer, void* stream);

/* Profiling hook: when non-NULL, every tcgen05 conv CTA of layer `layer` writes 8 clock64() stamps (start, setup done,
 * first operands landed, MMAs issued, accumulator ready, epilogue done, exit) to device_buffer[8*cta]. */
int svs_debug_set_trace(long long* device_buffer, int layer);

/* Debug / parity hook: copy an intermediate activation of the LAST svs_unet_forward call on this
 * workspace out as fp32 NCHW.  layer 0..5 = conv1..conv6 outputs, 6..10 = deconv1..deconv5 outputs. */
int svs_unet_read_activation(const svs_unet_plan* plan, int layer, int batch, const void* workspace,
                             float* out_nchw, void* stream);

/* Number of kernel launches svs_unet_forward enqueues for this plan/batch (bench bookkeeping). */
int svs_unet_launch_count(const svs_unet_plan* plan, int batch);

/* ------------------------------------------------------------------ host-buffer patch batches
 * Replaces the per-patch `.to(device)` -> model -> `.cpu()` round trip of reference inference.py:97-110 for callers
 * whose patches live in (pinned) HOST memory.  Step i uploads host_in[i] (batch x 512 x 128 float32, dense) into
 * device staging slot i % n_slots on stream_h2d, runs svs_unet_forward on stream_compute and downloads the result
 * into host_out[i] on stream_d2h; batch i+1's upload and batch i-1's download overlap batch i's kernels, ordered by
 * events the stream object owns.  Everything is enqueued; the call does not synchronise (wait on stream_d2h).
 *   host_in / host_out   HOST arrays of n_steps host pointers (page-locked for the copies to be asynchronous)
 *   dev_in / dev_out     HOST arrays of n_slots device pointers, batch*512*128 floats each
 *   workspace            one device workspace of svs_unet_workspace_bytes(plan, batch) (forwards are serialised)
 */
typedef struct svs_patch_stream svs_patch_stream;   /* opaque: 3 x n_slots events */
int svs_patch_stream_create(int n_slots, svs_patch_stream** out);
int svs_patch_stream_destroy(svs_patch_stream* ps);
int svs_patch_stream_run(svs_patch_stream* ps, const svs_unet_plan* plan, const float* const* host_in,
                         float* const* host_out, int n_steps, int batch, int flags, float* const* dev_in,
                         float* const* dev_out, void* workspace, size_t workspace_bytes, void* stream_h2d,
                         void* stream_compute, void* stream_d2h);

/* ------------------------------------------------------------------ training step (T1)
 * Replaces the autograd graph of reference train.py:274-299 (mask = model(mix) in train mode, L1 loss,
 * loss.backward()): train-mode forward (batch-statistic BatchNorm + running-stat update with momentum
 * 0.1 / unbiased variance, Dropout2d through explicit per-(sample, channel) keep masks), the masked-L1
 * loss of train.py:275-283 with crit = L1 (reference config.py:33,44), and the backward pass producing
 * every parameter gradient.  Parameters and gradients are the raw fp32 torch tensors of the state_dict
 * (torch layouts, no plan); torch.optim.Adam (reference model.py:116) consumes the gradients unchanged.
 * Arithmetic is fp32; all reductions have a fixed order (bit-reproducible, no atomics).
 */
typedef struct svs_train_layer {
  const float* weight;  const float* bias;     /* parameters (read)                                    */
  const float* bn_weight; const float* bn_bias;   /* NULL for deconv6                                  */
  float* bn_running_mean; float* bn_running_var;  /* updated in place by the forward when requested    */
  float* grad_weight; float* grad_bias;        /* written by the backward (overwritten, not accumulated) */
  float* grad_bn_weight; float* grad_bn_bias;
  const uint8_t* dropout_keep;                 /* device [batch][Cout] 0/1, or NULL (= keep all); only
                                                  deconv1..deconv5 (reference model.py:80-108)            */
} svs_train_layer;

/* Arithmetic of the step.  A train plan selects TF32 on the tensor cores (tcgen05 kind::tf32, fp32 accumulate) for the
 * forward, data-gradient and weight-gradient convolutions — what torch + cuDNN do by default for the reference's
 * fp32 model on a GPU (torch.backends.cudnn.allow_tf32) — while BatchNorm statistics, normalisation, the loss and every
 * reduction stay fp32.  plan == NULL runs the whole step in exact fp32 on CUDA cores (parity mode).  The plan owns
 * device scratch for repacked weights (rewritten by every forward call) and one internal side stream: every backward
 * call forks the weight gradients and the data-gradient weight packing onto it and joins it back into the caller's
 * stream before returning (also under CUDA-graph capture), so all of a call's work is ordered on the caller's stream
 * as usual.  The plan holds no parameters and may be shared by successive steps, not by concurrent ones. */
typedef struct svs_train_plan svs_train_plan;
int svs_unet_train_plan_create(void* stream, svs_train_plan** plan_out);
int svs_unet_train_plan_destroy(svs_train_plan* plan);

size_t svs_unet_train_workspace_bytes(int batch);

/* mix float32 dense (batch,1,512,128) -> mask float32 dense (batch,1,512,128).  The workspace (1024-byte aligned)
 * keeps everything the backward needs and must stay untouched until svs_unet_train_backward returns. */
int svs_unet_train_forward(const svs_train_plan* plan, const svs_train_layer layers[12], const float* mix, int batch,
                           int update_running_stats, float* mask_out, void* workspace, size_t workspace_bytes,
                           void* stream);

/* grad_mask = dLoss/dmask (dense, same shape as the mask).  Fills grad_* of all 12 layers.  `plan` must be the one
 * (or NULL) the forward ran with. */
int svs_unet_train_backward(const svs_train_plan* plan, const svs_train_layer layers[12], const float* mix,
                            const float* grad_mask, int batch, void* workspace, size_t workspace_bytes, void* stream);

/* The same backward restricted to layers [first_layer, last_layer] (0 = conv1 .. 11 = deconv6), run from last to first.
 * A data-parallel host calls it in descending segments and starts the NCCL all-reduce of a finished segment's
 * gradients while the next segment computes (reference train.py:298 has no data parallelism). */
int svs_unet_train_backward_layers(const svs_train_plan* plan, const svs_train_layer layers[12], const float* mix,
                                   const float* grad_mask, int batch, void* workspace, size_t workspace_bytes,
                                   int first_layer, int last_layer, void* stream);

/* The weight-gradient contraction on its own (the wgrad half of loss.backward() through nn.Conv2d /
 * nn.ConvTranspose2d, reference model.py:47-109), TF32 on tcgen05:
 *     grad_w[(m * l_c + n) * 25 + kh * 5 + kw] = sum_{b,y,x} S[b,y,x,s_coff+m] * L[b, 2y+kh-2, 2x+kw-2, l_coff+n]
 * S: fp32 NHWC [batch][gh][gw][s_pitch] (small grid; Conv2d: dL/dout, ConvTranspose2d: the layer input),
 * L: fp32 NHWC [batch][2gh][2gw][l_pitch] (Conv2d: the layer input, ConvTranspose2d: dL/dout); out-of-range taps
 * read zero.  Result in the torch layout (s_c, l_c, 5, 5).  Requirements: s_c % 32 == 0, l_c % 16 == 0, pitches and
 * offsets % 4 == 0, power-of-two grids; values should be TF32-representable (the tensor core truncates).
 * `partial`: device scratch of svs_conv_wgrad_partial_floats() floats (fixed-order split reduction). */
size_t svs_conv_wgrad_partial_floats(int gh, int gw, int batch, int s_c, int l_c);
int svs_conv_wgrad_tf32(const float* small, int s_pitch, int s_coff, int s_c, const float* large, int l_pitch,
                        int l_coff, int l_c, int gh, int gw, int batch, float* partial, size_t partial_floats,
                        float* grad_w, void* stream);

/* Fused loss of reference train.py:275-283 and its gradient w.r.t. the mask:
 *   L = mean|m*x - v| (+ mean|(1-m)*x - max(x - v, 0)| when two_term)      n = number of elements
 * loss_out: device float32 [3] = {total, vocal term, accompaniment term}; grad_mask_out may be NULL.
 * scratch: device float32 [SVS_L1_SCRATCH_FLOATS] owned by the caller (per-block partial sums of the two-stage,
 * fixed-order reduction; concurrent calls on different streams need different scratch buffers). */
#define SVS_L1_SCRATCH_FLOATS 2048
int svs_l1_masked_loss(const float* mask, const float* mix, const float* voc, int64_t n, int two_term,
                       float grad_scale, float* loss_out, float* grad_mask_out, float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVS_B200_H */
