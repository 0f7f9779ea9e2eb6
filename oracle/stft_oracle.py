"""CPU oracle: librosa-0.10.1-faithful STFT / magphase / iSTFT (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED at the librosa boundary: ``librosa==0.10.1`` (reference
``uv.lock:713-714``, ``requirements.txt:12``) is not vendored under
/root/reference and cannot be installed here, and the reference ships no golden
vectors for it.  The functions below restate the published librosa 0.10.1
algorithm as it is reached from the reference's call sites; each cites the call
site it follows.  ``tests/test_oracle_spectral.py`` cross-checks them against
``torch.stft`` / ``torch.istft`` (float64) as an independent second statement.

Semantics restated (librosa 0.10.1 ``core/spectrum.py``):

* ``stft``      window = scipy ``get_window('hann', n_fft, fftbins=True)`` (float64,
                periodic), ``center=True``, ``pad_mode='constant'`` (the 0.10 default),
                frames ``1 + len // hop``; ``scipy.fft.rfft`` of ``window * frames``
                evaluated in float64 (float64 window promotes the float32 audio), result
                cast to complex64, Fortran-ordered ``(1 + n_fft/2, T)``.
* ``magphase``  ``mag = abs(D)``; ``phase = D / (mag + [mag==0]) + [mag==0]`` with real and
                imaginary parts divided separately.
* ``istft``     ``n_fft = 2 (rows - 1)``; ``irfft`` on complex64 (float32 math in pocketfft),
                multiplied by the float64 window, overlap-added sequentially into a float32
                buffer (head frame handled separately because ``center=True``), divided by the
                float32 window-sum-of-squares envelope where it exceeds ``tiny(float32)``,
                output length ``hop * (T - 1)``.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.signal

N_FFT = 1024      # reference config.py:47  WINDOW_SIZE
HOP = 768         # reference config.py:48  HOP_SIZE
SR = 8192         # reference config.py:49  SAMPLE_RATE


def hann_periodic(n_fft: int = N_FFT) -> np.ndarray:
    """float64 periodic Hann, as librosa.filters.get_window('hann', n, fftbins=True)."""
    return scipy.signal.get_window("hann", n_fft, fftbins=True)


def n_frames(length: int, hop: int = HOP) -> int:
    """Frames produced by librosa.stft(center=True): 1 + len // hop (reference data.py:79)."""
    return 1 + length // hop


def stft(y: np.ndarray, n_fft: int = N_FFT, hop_length: int = HOP) -> np.ndarray:
    """``librosa.stft(y, n_fft=n_fft, hop_length=hop_length)`` as called at reference
    data.py:79 and data.py:100.  Returns complex64 ``(1 + n_fft//2, T)``, Fortran order."""
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError("oracle stft expects mono audio (len,)")
    if not np.issubdtype(y.dtype, np.floating):
        raise ValueError("audio must be floating point")
    window = hann_periodic(n_fft)                              # float64
    ypad = np.pad(y, (n_fft // 2, n_fft // 2), mode="constant")   # center=True, pad_mode="constant"
    t = 1 + (ypad.shape[0] - n_fft) // hop_length
    # frame t covers ypad[t*hop : t*hop + n_fft]  (librosa.util.frame, axis=-1 -> (n_fft, T))
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(t)[None, :]
    frames = ypad[idx]                                         # (n_fft, T), dtype of y
    spec = scipy.fft.rfft(window[:, None] * frames, axis=0)    # float64 math
    out = np.zeros((1 + n_fft // 2, t), dtype=np.complex64, order="F")
    out[...] = spec                                            # rounds to complex64
    return out


def magphase(d: np.ndarray):
    """``librosa.magphase(D)`` as called at reference data.py:80 / data.py:101."""
    mag = np.abs(d)
    zeros_to_ones = mag == 0
    mag_nonzero = mag + zeros_to_ones
    phase = np.empty_like(d, dtype=np.complex64 if d.dtype == np.complex64 else np.complex128)
    phase.real = d.real / mag_nonzero + zeros_to_ones
    phase.imag = d.imag / mag_nonzero
    return mag, phase


def to_spec(y_mix: np.ndarray, y_track: np.ndarray | None = None):
    """The ``to_spec`` body of reference data.py:78-105 for one song.

    Returns ``(spec, phase, norm)``: ``spec`` float32 ``(513, T)`` divided by the MIXTURE's
    max magnitude (``norm == 0 -> 1``), ``phase`` complex64 unit phasors."""
    spec_mix, _ = magphase(stft(y_mix))
    spec_mix = np.abs(spec_mix).astype(np.float32)
    norm = spec_mix.max()
    if norm == 0:
        norm = 1
    y = y_mix if y_track is None else y_track
    if len(y) > len(y_mix):                                    # data.py:97-98 length alignment
        y = y[: len(y_mix)]
    else:
        y = np.pad(y, (0, len(y_mix) - len(y)))
    spec, phase = magphase(stft(y))
    spec = np.abs(spec).astype(np.float32)
    spec /= norm
    return spec, phase, np.float32(norm)


def _overlap_add(y: np.ndarray, ytmp: np.ndarray, hop_length: int) -> None:
    """librosa ``__overlap_add`` (numba loop): sequential ``y[t*hop : t*hop+n_fft] += ytmp[:, t]``
    with the float64 frame cast into the float32 accumulator at each add."""
    n_fft = ytmp.shape[0]
    n = y.shape[-1]
    for frame in range(ytmp.shape[1]):
        sample = frame * hop_length
        if n_fft > n - sample:
            y[sample:] += ytmp[: n - sample, frame]
        else:
            y[sample: sample + n_fft] += ytmp[:, frame]


def window_sumsquare(n_frames_: int, hop_length: int = HOP, n_fft: int = N_FFT,
                     dtype=np.float32) -> np.ndarray:
    """librosa.filters.window_sumsquare('hann', ...) incl. the ``__window_ss_fill`` loop."""
    n = n_fft + hop_length * (n_frames_ - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = hann_periodic(n_fft) ** 2
    for i in range(n_frames_):
        sample = i * hop_length
        x[sample: min(n, sample + n_fft)] += win_sq[: max(0, min(n_fft, n - sample))]
    return x


def istft(stft_matrix: np.ndarray, win_length: int = N_FFT, hop_length: int = HOP) -> np.ndarray:
    """``librosa.istft(S, win_length=, hop_length=)`` as called at reference data.py:159."""
    n_fft = 2 * (stft_matrix.shape[-2] - 1)
    if win_length != n_fft:
        raise ValueError("oracle istft restates the win_length == n_fft case used by the reference")
    window = hann_periodic(win_length)[:, None]                # float64
    t = stft_matrix.shape[-1]
    rdtype = np.float32 if stft_matrix.dtype == np.complex64 else np.float64
    expected_len = n_fft + hop_length * (t - 1) - 2 * (n_fft // 2)
    y = np.zeros(expected_len, dtype=rdtype)

    def irfft(block):
        # pocketfft runs complex64 input in float32; complex128 in float64
        return scipy.fft.irfft(block, n=n_fft, axis=0)

    # center=True: frames that reach into the left padding are overlap-added in a head buffer
    start_frame = int(np.ceil((n_fft // 2) / hop_length))
    ytmp = window * irfft(stft_matrix[:, :start_frame])
    head_len = n_fft + hop_length * (start_frame - 1)
    head = np.zeros(head_len, dtype=rdtype)
    _overlap_add(head, ytmp, hop_length)
    if y.shape[-1] < head_len - n_fft // 2:
        y[:] = head[n_fft // 2: y.shape[-1] + n_fft // 2]
    else:
        y[: head_len - n_fft // 2] = head[n_fft // 2:]
    offset = start_frame * hop_length - n_fft // 2
    if t > start_frame:
        ytmp = window * irfft(stft_matrix[:, start_frame:])
        _overlap_add(y[offset:], ytmp, hop_length)

    win_sum = window_sumsquare(t, hop_length, n_fft, dtype=rdtype)[n_fft // 2:]
    if win_sum.shape[0] < y.shape[0]:
        win_sum = np.pad(win_sum, (0, y.shape[0] - win_sum.shape[0]))
    win_sum = win_sum[: y.shape[0]]
    nz = win_sum > np.finfo(win_sum.dtype).tiny
    y[nz] /= win_sum[nz]
    return y


def to_wave(mag: np.ndarray, phase: np.ndarray) -> np.ndarray:
    """The ``to_wave`` body of reference data.py:151-164 for one song (before sf.write)."""
    min_len = min(mag.shape[1], phase.shape[1])
    mag = mag[:, :min_len]
    phase = phase[:, :min_len]
    y = istft(mag * phase, win_length=N_FFT, hop_length=HOP)
    max_val = np.max(np.abs(y)) if y.size else 0.0
    if max_val > 0:
        y = y / max_val * 0.9
    return y
