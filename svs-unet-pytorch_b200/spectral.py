"""Spectral front / back end on the GPU: librosa-shaped calls + the ragged-batch device API.

librosa-shaped (drop-in for the call sites of reference data.py:79-81,100-102,159; numpy in/out):
    stft(y, n_fft=1024, hop_length=768)          -> complex64 (513, T), Fortran order
    magphase(D)                                  -> (float32 mag, complex64 unit phase)
    istft(S, win_length=1024, hop_length=768)    -> float32 (768 * (T - 1),)

device API (what the fused pipeline uses; everything stays in HBM):
    SongBatch.from_audio(list of arrays)  ->  .stft() / .normalize() / istft_batch(...)
Every function raises if the CUDA library is unavailable; nothing here computes on the CPU.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .config import HOP_SIZE, N_BINS, WINDOW_SIZE


def _check_geometry(n_fft, hop_length):
    if n_fft != WINDOW_SIZE or hop_length != HOP_SIZE:
        raise _lib.SvsError(f"the sm_100a kernels are specialised for n_fft={WINDOW_SIZE}, hop={HOP_SIZE} "
                            f"(reference config.py:47-48); got n_fft={n_fft}, hop={hop_length}")


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.SvsError("no CUDA device: svs-unet-pytorch_b200 has no CPU path")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


class SongBatch:
    """A ragged batch of mono songs resident on the device (concatenated samples + offset tables)."""

    _offset_cache: dict = {}

    def __init__(self, audio: torch.Tensor, lengths: list[int]):
        _lib.require_cuda(audio, "audio")
        if audio.dtype not in (torch.float32, torch.int16):
            raise _lib.SvsError(f"audio must be float32 samples or int16 PCM, got {audio.dtype}")
        self.audio = audio                                                # int16: the /32768 of librosa.load is fused into K1
        self.lengths = [int(n) for n in lengths]
        self.n_songs = len(self.lengths)
        self.frames = [1 + n // HOP_SIZE for n in self.lengths]           # librosa: 1 + len // hop
        self.wave_lengths = [HOP_SIZE * (t - 1) for t in self.frames]     # istft: hop * (T - 1)
        dev = audio.device
        # offset tables of the last few batch geometries stay on the device: building them costs three small
        # synchronous H2D copies, which stall a pipelined caller (SongStreamer) on every chunk
        key = (tuple(self.lengths), str(dev))
        hit = SongBatch._offset_cache.get(key)
        if hit is None:
            so = np.concatenate([[0], np.cumsum(self.lengths)]).astype(np.int64)
            fo = np.concatenate([[0], np.cumsum(self.frames)]).astype(np.int64)
            wo = np.concatenate([[0], np.cumsum(self.wave_lengths)]).astype(np.int64)
            hit = (so, fo, wo, torch.from_numpy(so).to(dev), torch.from_numpy(fo).to(dev), torch.from_numpy(wo).to(dev))
            if len(SongBatch._offset_cache) >= 8:
                SongBatch._offset_cache.pop(next(iter(SongBatch._offset_cache)))
            SongBatch._offset_cache[key] = hit
        so, fo, wo, self.sample_off, self.frame_off, self.wave_off = hit
        self.sample_off_host, self.frame_off_host, self.wave_off_host = so, fo, wo
        self.total_frames = int(fo[-1])
        self.total_wave = int(wo[-1])
        self.max_frames = max(self.frames)

    @classmethod
    def from_audio(cls, songs, device=None, pinned: torch.Tensor | None = None):
        """songs: list of 1-D float32 numpy arrays / CPU tensors (host) or CUDA tensors."""
        dev = _device(device)
        lengths = [int(len(s)) for s in songs]
        if all(isinstance(s, torch.Tensor) and s.is_cuda for s in songs):
            audio = torch.cat([s.float() for s in songs]) if len(songs) > 1 else songs[0].float().contiguous()
        else:
            host = pinned if pinned is not None else torch.empty(sum(lengths), dtype=torch.float32).pin_memory()
            off = 0
            for s in songs:
                host[off:off + len(s)] = torch.as_tensor(np.asarray(s, dtype=np.float32)) if not isinstance(s, torch.Tensor) else s
                off += len(s)
            audio = host[:off].to(dev, non_blocking=True)
        return cls(audio, lengths)

    def stft(self, want_phase: bool = True):
        """-> (mag [F,513] f32, phase [F,513,2] f32 | None, song_max [n_songs] f32); F = total frames."""
        return _lib.stft_mag_phase_raw(self.audio, self.sample_off, self.frame_off, self.n_songs,
                                       self.max_frames, self.total_frames, want_phase=want_phase)

    def stft_complex(self):
        return _lib.stft_complex_raw(self.audio, self.sample_off, self.frame_off, self.n_songs,
                                     self.max_frames, self.total_frames)

    def normalize(self, mag: torch.Tensor, norm: torch.Tensor):
        """spec /= norm per song (reference data.py:105); in place."""
        return _lib.spec_normalize_raw(mag, self.frame_off, norm, self.n_songs)

    def istft(self, mag: torch.Tensor, phase: torch.Tensor, peak_normalize: bool = False, pcm16: bool = False):
        """-> (wave [total_wave] f32, song_peak [n_songs] f32).  Optional 0.9 peak normalisation
        (reference data.py:162-164).  ``pcm16=True`` returns int16 PCM instead: the normalisation fused with the
        PCM_16 quantiser of ``sf.write`` (data.py:166), i.e. the bytes of the restored .wav file."""
        wave, peak = _lib.istft_ola_raw(mag, phase, self.frame_off, self.wave_off, self.n_songs,
                                        self.max_frames, self.total_wave)
        if pcm16:
            if not peak_normalize:
                raise _lib.SvsError("pcm16 output implies the 0.9 peak normalisation of data.py:162-166")
            if self.total_wave > 0:
                wave = _lib.wave_peak_normalize_pcm16_raw(wave, self.wave_off, peak, self.n_songs, 0.9)
            else:
                wave = wave.to(torch.int16)
        elif peak_normalize and self.total_wave > 0:
            _lib.wave_peak_normalize_raw(wave, self.wave_off, peak, self.n_songs, 0.9)
        return wave, peak

    def song_spec(self, mag: torch.Tensor, s: int) -> torch.Tensor:
        """(513, T_s) view of song s — the Fortran-ordered array librosa would return."""
        a, b = int(self.frame_off_host[s]), int(self.frame_off_host[s + 1])
        return mag[a:b].transpose(0, 1)

    def song_wave(self, wave: torch.Tensor, s: int) -> torch.Tensor:
        a = int(self.wave_off_host[s])
        return wave[a:a + self.wave_lengths[s]]


# ---------------------------------------------------------------------------------------------
# librosa-shaped single-song calls (numpy in / numpy out)

def stft(y, n_fft: int = WINDOW_SIZE, hop_length: int = HOP_SIZE) -> np.ndarray:
    _check_geometry(n_fft, hop_length)
    y = np.ascontiguousarray(y, dtype=np.float32)
    if y.ndim != 1:
        raise _lib.SvsError("stft expects mono audio of shape (len,)")
    batch = SongBatch.from_audio([y])
    spec = batch.stft_complex()                                     # [T, 513, 2]
    d = torch.view_as_complex(spec).cpu().numpy()                   # (T, 513) C-order
    return d.T                                                      # (513, T) Fortran-ordered view


def magphase(d):
    d = np.asarray(d)
    if d.dtype != np.complex64:
        d = d.astype(np.complex64)
    f_order = d.flags["F_CONTIGUOUS"] and not d.flags["C_CONTIGUOUS"]
    flat = np.ascontiguousarray(d.T if f_order else d)
    t = torch.view_as_real(torch.from_numpy(flat)).to(_device())
    mag, phase = _lib.magphase_raw(t)
    mag = mag.cpu().numpy()
    phase = torch.view_as_complex(phase).cpu().numpy()
    return (mag.T, phase.T) if f_order else (mag, phase)


def istft(stft_matrix, win_length: int = WINDOW_SIZE, hop_length: int = HOP_SIZE) -> np.ndarray:
    """``librosa.istft(mag * phase, win_length=, hop_length=)`` (reference data.py:159)."""
    s = np.asarray(stft_matrix)
    if s.shape[0] != N_BINS:
        raise _lib.SvsError(f"istft expects ({N_BINS}, T); got {s.shape}")
    _check_geometry(2 * (s.shape[0] - 1), hop_length)
    if win_length != WINDOW_SIZE:
        raise _lib.SvsError("win_length must equal n_fft = 1024")
    t = s.shape[1]
    dev = _device()
    st = np.ascontiguousarray(s.T.astype(np.complex64))             # [T, 513]
    phase = torch.view_as_real(torch.from_numpy(st)).to(dev)        # treated as (1 * complex)
    mag = torch.ones((t, N_BINS), dtype=torch.float32, device=dev)
    return istft_mag_phase(mag, phase)


def istft_mag_phase(mag_tf: torch.Tensor, phase_tf: torch.Tensor) -> np.ndarray:
    """One song: mag [T,513], phase [T,513,2] device tensors -> float32 numpy waveform."""
    t = mag_tf.shape[0]
    batch = SongBatch(torch.zeros(1, dtype=torch.float32, device=mag_tf.device), [HOP_SIZE * (t - 1)])
    assert batch.frames[0] == t
    wave, _ = batch.istft(mag_tf.contiguous(), phase_tf.contiguous())
    return wave.cpu().numpy()
