"""GPU decode side: PCM (int16 / float32, any channel count) -> mono float32 at the model rate.

The arithmetic of ``librosa.load(path, sr=8192, mono=True)`` (reference data.py:78,94) after the file has been read:
sample conversion, channel mean and rational resampling, fused in ``svs_resample_poly``.  The filter is the one
``scipy.signal.resample_poly`` designs (Kaiser beta 5 windowed sinc, half length 10 * max(up, down)); librosa's
default ``soxr_hq`` is a different filter whose coefficients are not reproducible here, so samples differ from the
reference's at the level any two high-quality resamplers differ ("needs its own oracle").
No CPU path: the design of the ~2e5 filter taps is host numpy, the filtering is the CUDA kernel."""
from __future__ import annotations

from math import gcd

import numpy as np
import torch

from . import _lib

_FILTERS: dict = {}


def design(up: int, down: int):
    """scipy.signal.resample_poly's filter and padding for (up, down): returns (h float64 [L], pre_pad, pre_remove,
    taps per phase).  firwin(2 * half_len + 1, 1 / max_rate, window=('kaiser', 5.0)) * up, restated in numpy."""
    max_rate = max(up, down)
    half_len = 10 * max_rate
    n = 2 * half_len + 1
    m = np.arange(n, dtype=np.float64) - half_len
    fc = 1.0 / max_rate
    h = fc * np.sinc(fc * m) * np.kaiser(n, 5.0)
    h /= h.sum()                                                       # firwin: unit gain at DC
    h *= up
    pre_pad = down - half_len % down
    pre_remove = (half_len + pre_pad) // down
    taps = -(-n // up)
    return h, int(pre_pad), int(pre_remove), int(taps)


def _filter(up: int, down: int, device):
    key = (up, down, str(device))
    hit = _FILTERS.get(key)
    if hit is None:
        h, pre_pad, pre_remove, taps = design(up, down)
        hp = np.zeros((up, taps), dtype=np.float32)
        idx = np.arange(len(h))
        hp[idx % up, idx // up] = h.astype(np.float32)
        hit = (torch.from_numpy(hp).to(device), pre_pad, pre_remove, taps)
        if len(_FILTERS) >= 4:
            _FILTERS.pop(next(iter(_FILTERS)))
        _FILTERS[key] = hit
    return hit


def out_length(n_in: int, up: int, down: int) -> int:
    return -(-n_in * up // down)


def resample_songs(pcm: torch.Tensor, frames_per_song, channels: int, in_sr: int, out_sr: int) -> tuple[torch.Tensor, list[int]]:
    """pcm: device tensor of interleaved frames (int16 or float32), songs back to back.  Returns (mono float32 device
    tensor with the songs back to back at ``out_sr``, samples per song)."""
    _lib.require_cuda(pcm, "pcm")
    if pcm.dtype not in (torch.int16, torch.float32):
        raise _lib.SvsError(f"pcm must be int16 or float32, got {pcm.dtype}")
    _lib.check_device(pcm.device)
    g = gcd(int(in_sr), int(out_sr))
    up, down = int(out_sr) // g, int(in_sr) // g
    hp, pre_pad, pre_remove, taps = _filter(up, down, pcm.device)
    frames = [int(n) for n in frames_per_song]
    outs = [out_length(n, up, down) for n in frames]
    in_off = torch.from_numpy(np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)).to(pcm.device)
    out_off = torch.from_numpy(np.concatenate([[0], np.cumsum(outs)]).astype(np.int64)).to(pcm.device)
    out = torch.empty(sum(outs), dtype=torch.float32, device=pcm.device)
    if sum(outs) == 0:
        return out, outs
    with torch.cuda.device(pcm.device):
        _lib.check(_lib.load().svs_resample_poly(pcm.data_ptr(), 1 if pcm.dtype == torch.int16 else 0, int(channels),
                                                 in_off.data_ptr(), out_off.data_ptr(), len(frames), max(outs), up, down,
                                                 pre_pad, pre_remove, hp.data_ptr(), taps, out.data_ptr(),
                                                 _lib.stream_ptr(pcm.device)), "svs_resample_poly")
    return out, outs


def resample(x: np.ndarray, in_sr: int, out_sr: int) -> np.ndarray:
    """numpy (n,) or (n, channels), float32 or int16 -> mono float32 (ceil(n * out_sr / in_sr),) through the GPU."""
    x = np.ascontiguousarray(x)
    if x.dtype not in (np.int16, np.float32):
        x = x.astype(np.float32)
    channels = 1 if x.ndim == 1 else x.shape[1]
    dev = torch.device("cuda", torch.cuda.current_device())
    y, _ = resample_songs(torch.from_numpy(x.reshape(-1)).to(dev), [x.shape[0]], channels, in_sr, out_sr)
    return y.cpu().numpy()
