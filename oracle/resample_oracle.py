"""CPU oracle for the decode-side resampler (TEST INFRASTRUCTURE ONLY — never imported by the product package).

Reference call site: ``librosa.load(path, sr=SAMPLE_RATE, mono=True)`` at data.py:78,94, i.e. soundfile decode ->
``librosa.to_mono`` (channel mean) -> ``librosa.resample(res_type='soxr_hq')``.  soxr is absent from this image and
its filter coefficients are not published as a formula: PARITY UNPINNED at that boundary.  The oracle is therefore the
published algorithm the kernel restates — ``scipy.signal.resample_poly`` (Kaiser beta 5 polyphase FIR) in float64 —
applied after the same sample conversion and channel mean."""
from __future__ import annotations

from math import gcd

import numpy as np
import scipy.signal


def load_like(x: np.ndarray, in_sr: int, out_sr: int) -> np.ndarray:
    """x: (n,) or (n, channels) int16 / float -> mono float64 at out_sr."""
    x = np.asarray(x)
    y = x.astype(np.float64) / 32768.0 if x.dtype == np.int16 else x.astype(np.float64)
    if y.ndim == 2:
        y = y.mean(axis=1)                                             # librosa.to_mono
    if in_sr == out_sr:
        return y
    g = gcd(int(in_sr), int(out_sr))
    return scipy.signal.resample_poly(y, out_sr // g, in_sr // g)
