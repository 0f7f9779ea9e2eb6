"""ctypes binding of libsvs_b200.so (C ABI in include/svs_b200.h) + torch custom-op registration.

This module is the only place that talks to the shared library.  It never falls back: if the
library is missing (or cannot be built because nvcc is absent) every entry point raises."""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvs_b200.so")

SVS_PRECISION = {"fp32": 0, "bf16": 1, "tf32": 2}
FLAG_APPLY_MASK = 1
FLAG_INVERT = 2

N_FFT, HOP, N_BINS, PATCH_BINS, PATCH_FRAMES = 1024, 768, 513, 512, 128
ABI_VERSION = 3
L1_SCRATCH_FLOATS = 2048


class ConvParams(Structure):
    _fields_ = [("weight", c_void_p), ("bias", c_void_p), ("bn_weight", c_void_p),
                ("bn_bias", c_void_p), ("bn_mean", c_void_p), ("bn_var", c_void_p)]


class PatchView(Structure):
    _fields_ = [("base", c_void_p), ("patch_off", c_void_p), ("stride_b", c_int64),
                ("stride_f", c_int64), ("stride_t", c_int64)]


class TrainLayer(Structure):
    _fields_ = [("weight", c_void_p), ("bias", c_void_p), ("bn_weight", c_void_p), ("bn_bias", c_void_p),
                ("bn_running_mean", c_void_p), ("bn_running_var", c_void_p),
                ("grad_weight", c_void_p), ("grad_bias", c_void_p),
                ("grad_bn_weight", c_void_p), ("grad_bn_bias", c_void_p), ("dropout_keep", c_void_p)]


class SvsError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()

# every symbol include/svs_b200.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ("svs_version", c_int, []),
    ("svs_last_error", c_char_p, []),
    ("svs_device_check", c_int, [c_int]),
    ("svs_resample_poly", c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int64,
                                  c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    ("svs_stft_mag_phase", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("svs_stft_mag_phase_pcm16", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("svs_stft_complex", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    ("svs_magphase", c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    ("svs_spec_normalize", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    ("svs_istft_ola", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    ("svs_wave_peak_normalize", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float, c_void_p]),
    ("svs_wave_peak_normalize_pcm16", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float, c_void_p, c_void_p]),
    ("svs_patches_gather", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    ("svs_patches_scatter", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    ("svs_unet_plan_create", c_int, [POINTER(ConvParams), c_int, c_void_p, POINTER(c_void_p)]),
    ("svs_unet_plan_destroy", c_int, [c_void_p]),
    ("svs_unet_plan_precision", c_int, [c_void_p]),
    ("svs_unet_workspace_bytes", c_size_t, [c_void_p, c_int]),
    ("svs_unet_forward", c_int, [c_void_p, POINTER(PatchView), POINTER(PatchView), c_void_p, c_int, c_int,
                                 c_void_p, c_size_t, c_void_p]),
    ("svs_unet_forward_layers", c_int, [c_void_p, POINTER(PatchView), POINTER(PatchView), c_void_p, c_int, c_int,
                                        c_void_p, c_size_t, c_int, c_int, c_void_p]),
    ("svs_debug_set_trace", c_int, [c_void_p, c_int]),
    ("svs_unet_read_activation", c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    ("svs_unet_launch_count", c_int, [c_void_p, c_int]),
    ("svs_patch_stream_create", c_int, [c_int, POINTER(c_void_p)]),
    ("svs_patch_stream_destroy", c_int, [c_void_p]),
    ("svs_patch_stream_run", c_int, [c_void_p, c_void_p, POINTER(c_void_p), POINTER(c_void_p), c_int, c_int, c_int,
                                     POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_size_t, c_void_p, c_void_p,
                                     c_void_p]),
    ("svs_unet_train_plan_create", c_int, [c_void_p, POINTER(c_void_p)]),
    ("svs_unet_train_plan_destroy", c_int, [c_void_p]),
    ("svs_unet_train_workspace_bytes", c_size_t, [c_int]),
    ("svs_unet_train_forward", c_int, [c_void_p, POINTER(TrainLayer), c_void_p, c_int, c_int, c_void_p, c_void_p,
                                       c_size_t, c_void_p]),
    ("svs_unet_train_backward", c_int, [c_void_p, POINTER(TrainLayer), c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                                        c_void_p]),
    ("svs_unet_train_backward_layers", c_int, [c_void_p, POINTER(TrainLayer), c_void_p, c_void_p, c_int, c_void_p,
                                               c_size_t, c_int, c_int, c_void_p]),
    ("svs_conv_wgrad_partial_floats", c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    ("svs_conv_wgrad_tf32", c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p, c_size_t, c_void_p, c_void_p]),
    ("svs_l1_masked_loss", c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]


def load(build_if_missing: bool = True):
    """Returns the loaded CDLL; builds it in-tree with nvcc first if needed.  Raises if impossible."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_missing:
            from . import build as _build
            try:
                _build.build()
            except Exception as e:  # no fallback: surface the failure
                if not os.path.exists(LIB_PATH):
                    raise SvsError(f"libsvs_b200.so is missing and could not be built: {e}") from e
        if not os.path.exists(LIB_PATH):
            raise SvsError(f"{LIB_PATH} not found — run `python __graft_entry__.py` (build()) first; "
                           "there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, restype, argtypes in _SIGNATURES:
            fn = getattr(lib, name)          # AttributeError = ABI mismatch, fail loudly
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.svs_version() != ABI_VERSION:
            raise SvsError("libsvs_b200.so ABI version mismatch")
        _lib = lib
        return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().svs_last_error()
        raise SvsError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise SvsError(f"{name} must be a CUDA tensor: svs-unet-pytorch_b200 has no CPU path "
                       "(the CPU restatement lives in oracle/ and is test infrastructure only)")
    if dtype is not None and t.dtype != dtype:
        raise SvsError(f"{name} must have dtype {dtype}, got {t.dtype}")


_checked_devices = set()


def check_device(device: torch.device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _checked_devices:
        check(load().svs_device_check(idx), "svs_device_check")
        _checked_devices.add(idx)


# ---------------------------------------------------------------------------------------------
# raw wrappers (device tensors in, device tensors out)

def stft_mag_phase_raw(audio, sample_off, frame_off, n_songs, max_frames, total_frames, want_phase=True,
                       want_max=True):
    require_cuda(audio, "audio")
    if audio.dtype not in (torch.float32, torch.int16):
        raise SvsError(f"audio must be float32 or int16 PCM, got {audio.dtype}")
    check_device(audio.device)
    dev = audio.device
    mag = torch.empty((total_frames, N_BINS), dtype=torch.float32, device=dev)
    phase = torch.empty((total_frames, N_BINS, 2), dtype=torch.float32, device=dev) if want_phase else None
    smax = torch.empty((n_songs,), dtype=torch.float32, device=dev) if want_max else None
    fn = load().svs_stft_mag_phase if audio.dtype == torch.float32 else load().svs_stft_mag_phase_pcm16
    with torch.cuda.device(dev):
        check(fn(audio.data_ptr(), sample_off.data_ptr(), frame_off.data_ptr(), n_songs,
                                        max_frames, mag.data_ptr(),
                                        phase.data_ptr() if phase is not None else None,
                                        smax.data_ptr() if smax is not None else None, stream_ptr(dev)),
              "svs_stft_mag_phase")
    return mag, phase, smax


def stft_complex_raw(audio, sample_off, frame_off, n_songs, max_frames, total_frames):
    require_cuda(audio, "audio", torch.float32)
    check_device(audio.device)
    spec = torch.empty((total_frames, N_BINS, 2), dtype=torch.float32, device=audio.device)
    with torch.cuda.device(audio.device):
        check(load().svs_stft_complex(audio.data_ptr(), sample_off.data_ptr(), frame_off.data_ptr(), n_songs,
                                      max_frames, spec.data_ptr(), stream_ptr(audio.device)), "svs_stft_complex")
    return spec


def magphase_raw(spec):
    require_cuda(spec, "spec", torch.float32)
    n = spec.numel() // 2
    mag = torch.empty(spec.shape[:-1], dtype=torch.float32, device=spec.device)
    phase = torch.empty_like(spec)
    with torch.cuda.device(spec.device):
        check(load().svs_magphase(spec.data_ptr(), n, mag.data_ptr(), phase.data_ptr(), stream_ptr(spec.device)),
              "svs_magphase")
    return mag, phase


def spec_normalize_raw(mag, frame_off, norm, n_songs):
    require_cuda(mag, "mag", torch.float32)
    with torch.cuda.device(mag.device):
        check(load().svs_spec_normalize(mag.data_ptr(), frame_off.data_ptr(), norm.data_ptr(), n_songs,
                                        mag.shape[0], stream_ptr(mag.device)), "svs_spec_normalize")
    return mag


def istft_ola_raw(mag, phase, frame_off, wave_off, n_songs, max_frames, total_samples, want_peak=True):
    require_cuda(mag, "mag", torch.float32)
    require_cuda(phase, "phase", torch.float32)
    check_device(mag.device)
    dev = mag.device
    # single-frame songs have hop * (T - 1) = 0 output samples (librosa returns an empty array): a batch of them has an
    # empty waveform, whose data_ptr() is null -- pass the one-element backing store instead
    base = torch.empty((max(total_samples, 1),), dtype=torch.float32, device=dev)
    wave = base[:total_samples]
    peak = torch.empty((n_songs,), dtype=torch.float32, device=dev) if want_peak else None
    with torch.cuda.device(dev):
        check(load().svs_istft_ola(mag.data_ptr(), phase.data_ptr(), frame_off.data_ptr(), wave_off.data_ptr(),
                                   n_songs, max_frames, base.data_ptr(),
                                   peak.data_ptr() if peak is not None else None, stream_ptr(dev)),
              "svs_istft_ola")
    return wave, peak


def wave_peak_normalize_raw(wave, wave_off, peak, n_songs, target=0.9):
    with torch.cuda.device(wave.device):
        check(load().svs_wave_peak_normalize(wave.data_ptr(), wave_off.data_ptr(), peak.data_ptr(), n_songs,
                                             wave.numel(), target, stream_ptr(wave.device)),
              "svs_wave_peak_normalize")
    return wave


def wave_peak_normalize_pcm16_raw(wave, wave_off, peak, n_songs, target=0.9, out=None):
    """float32 waveforms -> int16 PCM with the 0.9 / peak normalisation fused (reference data.py:162-166)."""
    if out is None:
        out = torch.empty(wave.shape, dtype=torch.int16, device=wave.device)
    with torch.cuda.device(wave.device):
        check(load().svs_wave_peak_normalize_pcm16(wave.data_ptr(), wave_off.data_ptr(), peak.data_ptr(), n_songs,
                                                   wave.numel(), target, out.data_ptr(), stream_ptr(wave.device)),
              "svs_wave_peak_normalize_pcm16")
    return out


def patches_gather_raw(spec, patch_off, in_frames, norm, out=None):
    """Frame-major spectrogram [frames][513] -> dense patches [n][1][512][128] (reference inference.py:74-97), with
    the per-patch norm (data.py:85,105) folded in when given."""
    require_cuda(spec, "spec", torch.float32)
    n = int(patch_off.numel())
    if out is None:
        out = torch.empty((n, 1, 512, 128), dtype=torch.float32, device=spec.device)
    with torch.cuda.device(spec.device):
        check(load().svs_patches_gather(spec.data_ptr(), patch_off.data_ptr(),
                                        in_frames.data_ptr() if in_frames is not None else None,
                                        norm.data_ptr() if norm is not None else None, out.data_ptr(), n,
                                        stream_ptr(spec.device)), "svs_patches_gather")
    return out


def patches_scatter_raw(patches, patch_off, in_frames, spec, dc_zero=True):
    """Dense patches -> frame-major spectrogram (crop + concatenate + DC row of reference inference.py:110-127)."""
    require_cuda(patches, "patches", torch.float32)
    require_cuda(spec, "spec", torch.float32)
    n = int(patch_off.numel())
    with torch.cuda.device(spec.device):
        check(load().svs_patches_scatter(patches.data_ptr(), patch_off.data_ptr(),
                                         in_frames.data_ptr() if in_frames is not None else None, spec.data_ptr(), n,
                                         1 if dc_zero else 0, stream_ptr(spec.device)), "svs_patches_scatter")
    return spec


def conv_wgrad_tf32(small: torch.Tensor, large: torch.Tensor, s_coff: int = 0, s_c: int | None = None,
                    l_coff: int = 0, l_c: int | None = None) -> torch.Tensor:
    """Weight gradient of a 5x5 stride-2 Conv2d / ConvTranspose2d on tcgen05 (svs_conv_wgrad_tf32): ``small`` fp32 NHWC
    (B, gh, gw, Cs), ``large`` fp32 NHWC (B, 2gh, 2gw, Cl) -> (s_c, l_c, 5, 5)."""
    require_cuda(small, "small", torch.float32)
    require_cuda(large, "large", torch.float32)
    check_device(small.device)
    b, gh, gw, sp = small.shape
    lp = large.shape[3]
    s_c = sp - s_coff if s_c is None else s_c
    l_c = lp - l_coff if l_c is None else l_c
    n = load().svs_conv_wgrad_partial_floats(gh, gw, b, s_c, l_c)
    partial = torch.empty(n, dtype=torch.float32, device=small.device)
    out = torch.empty((s_c, l_c, 5, 5), dtype=torch.float32, device=small.device)
    with torch.cuda.device(small.device):
        check(load().svs_conv_wgrad_tf32(small.data_ptr(), sp, s_coff, s_c, large.data_ptr(), lp, l_coff, l_c, gh, gw, b,
                                         partial.data_ptr(), n, out.data_ptr(), stream_ptr(small.device)),
              "svs_conv_wgrad_tf32")
    return out


class TrainPlan:
    """svs_train_plan: TF32 tensor-core arithmetic for the training step (device scratch for repacked weights)."""

    def __init__(self, device):
        self.device = torch.device(device)
        check_device(self.device)
        handle = c_void_p()
        with torch.cuda.device(self.device):
            check(load().svs_unet_train_plan_create(stream_ptr(self.device), byref(handle)), "svs_unet_train_plan_create")
            torch.cuda.current_stream(self.device).synchronize()
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                load().svs_unet_train_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# UNet plans

class UNetPlan:
    """Owns one svs_unet_plan (BatchNorm-folded, repacked weights) and its activation workspace."""

    LAYERS = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6",
              "deconv1", "deconv2", "deconv3", "deconv4", "deconv5", "deconv6"]

    def __init__(self, state_dict: dict, precision: str = "bf16"):
        if precision not in SVS_PRECISION:
            raise SvsError(f"precision must be one of {list(SVS_PRECISION)}")
        lib = load()
        arr = (ConvParams * 12)()
        keep = []

        def ptr(key):
            t = state_dict[key]
            require_cuda(t, key)
            t = t.detach().to(torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        for i, name in enumerate(self.LAYERS):
            enc = i < 6
            arr[i].weight = ptr(f"{name}.0.weight" if enc else f"{name}.weight")
            arr[i].bias = ptr(f"{name}.0.bias" if enc else f"{name}.bias")
            bn = f"{name}.1" if enc else f"{name}_BAD.0"
            if name != "deconv6":
                arr[i].bn_weight = ptr(bn + ".weight")
                arr[i].bn_bias = ptr(bn + ".bias")
                arr[i].bn_mean = ptr(bn + ".running_mean")
                arr[i].bn_var = ptr(bn + ".running_var")
        self.device = keep[0].device
        check_device(self.device)
        self.precision = precision
        handle = c_void_p()
        with torch.cuda.device(self.device):
            check(lib.svs_unet_plan_create(arr, SVS_PRECISION[precision], stream_ptr(self.device), byref(handle)),
                  "svs_unet_plan_create")
            torch.cuda.current_stream(self.device).synchronize()
        self.handle = handle
        self._ws = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                load().svs_unet_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def workspace(self, batch: int) -> torch.Tensor:
        """Activation workspace for `batch` patches on the CURRENT stream.  svs_unet_forward writes every concat
        buffer into it, so two forwards in flight on different streams must not share one: the cache is keyed by
        (stream, batch).  A stream keeps the workspaces of its last few batch sizes (a corpus whose last batch is
        ragged alternates between two sizes: re-allocating on every switch cost the 19-song shard of the 8-GPU run
        two allocator round trips per pass).  A workspace is only ever used on the stream it was allocated under, so
        dropping it hands the block back to the caching allocator's pool of that same stream and any re-use is
        stream-ordered after the kernels that still read it."""
        key = (torch.cuda.current_stream(self.device).cuda_stream, batch)
        hit = self._ws.pop(key, None)
        if hit is None:
            nbytes = load().svs_unet_workspace_bytes(self.handle, batch)
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-raw.data_ptr()) % 1024
            hit = raw[off:off + nbytes]
            per_stream = [k for k in self._ws if k[0] == key[0]]
            if len(per_stream) >= 3:                                     # keep at most three more sizes per stream
                self._ws.pop(per_stream[0])
            while len(self._ws) >= 32:                                   # and bound the whole cache
                self._ws.pop(next(iter(self._ws)))
        self._ws[key] = hit                                              # most recently used last
        return hit

    def launch_count(self, batch: int) -> int:
        return load().svs_unet_launch_count(self.handle, batch)

    def forward_views(self, in_view: PatchView, out_view: PatchView, in_frames, batch: int, flags: int,
                      first_layer: int = 0, last_layer: int = 11):
        ws = self.workspace(batch)
        with torch.cuda.device(self.device):
            check(load().svs_unet_forward_layers(self.handle, byref(in_view), byref(out_view),
                                                 in_frames.data_ptr() if in_frames is not None else None, batch,
                                                 flags, ws.data_ptr(), ws.numel(), first_layer, last_layer,
                                                 stream_ptr(self.device)),
                  "svs_unet_forward")

    def forward_dense(self, mix: torch.Tensor, flags: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
        """mix float32 CUDA (B,1,512,128) contiguous -> float32 (B,1,512,128)."""
        require_cuda(mix, "mix", torch.float32)
        if mix.dim() != 4 or tuple(mix.shape[1:]) != (1, PATCH_BINS, PATCH_FRAMES):
            raise SvsError(f"mix must have shape (B, 1, 512, 128); got {tuple(mix.shape)} — other geometries "
                           "of reference config.py are not supported by the sm_100a kernels")
        mix = mix.contiguous()
        if out is None:
            out = torch.empty_like(mix)
        b = mix.shape[0]
        iv = PatchView(mix.data_ptr(), None, PATCH_BINS * PATCH_FRAMES, PATCH_FRAMES, 1)
        ov = PatchView(out.data_ptr(), None, PATCH_BINS * PATCH_FRAMES, PATCH_FRAMES, 1)
        self.forward_views(iv, ov, None, b, flags)
        return out

    def read_activation(self, layer: int, batch: int) -> torch.Tensor:
        geom = [(16, 256, 64), (32, 128, 32), (64, 64, 16), (128, 32, 8), (256, 16, 4), (512, 8, 2),
                (256, 16, 4), (128, 32, 8), (64, 64, 16), (32, 128, 32), (16, 256, 64)][layer]
        out = torch.empty((batch,) + geom, dtype=torch.float32, device=self.device)
        ws = self.workspace(batch)
        with torch.cuda.device(self.device):
            check(load().svs_unet_read_activation(self.handle, layer, batch, ws.data_ptr(), out.data_ptr(),
                                                  stream_ptr(self.device)), "svs_unet_read_activation")
        return out


# ---------------------------------------------------------------------------------------------
# torch custom ops (the reference-facing operator surface)

_PLANS: dict[int, UNetPlan] = {}
_next_plan_id = [1]


def register_plan(plan: UNetPlan) -> int:
    pid = _next_plan_id[0]
    _next_plan_id[0] += 1
    _PLANS[pid] = plan
    return pid


def release_plan(pid: int):
    _PLANS.pop(pid, None)


@torch.library.custom_op("svs_b200::unet_forward", mutates_args=())
def unet_forward_op(mix: torch.Tensor, plan_id: int, flags: int) -> torch.Tensor:
    plan = _PLANS.get(plan_id)
    if plan is None:
        raise SvsError(f"unknown UNet plan id {plan_id}")
    return plan.forward_dense(mix, flags)


@unet_forward_op.register_fake
def _(mix, plan_id, flags):
    return torch.empty_like(mix)


@torch.library.custom_op("svs_b200::stft_mag_phase", mutates_args=())
def stft_mag_phase_op(audio: torch.Tensor, sample_off: torch.Tensor, frame_off: torch.Tensor, n_songs: int,
                      max_frames: int, total_frames: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    mag, phase, smax = stft_mag_phase_raw(audio, sample_off, frame_off, n_songs, max_frames, total_frames)
    return mag, phase, smax


@stft_mag_phase_op.register_fake
def _(audio, sample_off, frame_off, n_songs, max_frames, total_frames):
    return (audio.new_empty((total_frames, N_BINS)), audio.new_empty((total_frames, N_BINS, 2)),
            audio.new_empty((n_songs,)))


@torch.library.custom_op("svs_b200::istft_ola", mutates_args=())
def istft_ola_op(mag: torch.Tensor, phase: torch.Tensor, frame_off: torch.Tensor, wave_off: torch.Tensor,
                 n_songs: int, max_frames: int, total_samples: int) -> tuple[torch.Tensor, torch.Tensor]:
    wave, peak = istft_ola_raw(mag, phase, frame_off, wave_off, n_songs, max_frames, total_samples)
    return wave, peak


@istft_ola_op.register_fake
def _(mag, phase, frame_off, wave_off, n_songs, max_frames, total_samples):
    return mag.new_empty((total_samples,)), mag.new_empty((n_songs,))
