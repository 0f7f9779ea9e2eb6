"""Synthetic workloads of BASELINE.json's configs; no dataset is available."""
from __future__ import annotations

import numpy as np

from .config import SAMPLE_RATE


def synth_song(seconds: float = 30.0, seed: int = 1234, sr: int = SAMPLE_RATE):
    """One synthetic mono mixture with its stems: ``(mixture, vocal, accomp)`` float32 ``(len,)``.

    vocal  = 6-harmonic tone stack, f0 gliding 110 -> 440 Hz, 4 Hz tremolo, peak 0.3
    accomp = 0.1 * N(0,1) + three fixed sines (196 / 247 / 294 Hz, amplitude 0.1)"""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    f0 = 110.0 * (440.0 / 110.0) ** (t / max(t[-1], 1e-9))
    ph = 2.0 * np.pi * np.cumsum(f0) / sr
    vocal = sum(np.sin(h * ph) / h for h in range(1, 7))
    vocal *= 0.5 * (1.0 + np.sin(2.0 * np.pi * 4.0 * t))
    vocal *= 0.3 / np.max(np.abs(vocal))
    accomp = 0.1 * rng.standard_normal(n)
    for f in (196.0, 247.0, 294.0):
        accomp += 0.1 * np.sin(2.0 * np.pi * f * t)
    vocal = vocal.astype(np.float32)
    accomp = accomp.astype(np.float32)
    return (vocal + accomp).astype(np.float32), vocal, accomp


def synth_patches(batch: int = 64, seed: int = 0):
    """Config 2: ``torch.manual_seed(seed); torch.rand(batch, 1, 512, 128)`` (U[0,1) like spec/norm)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 1, 512, 128, generator=g)


def sdr_db(ref: np.ndarray, est: np.ndarray) -> float:
    """Plain SDR after a least-squares gain (absolute scale is discarded by the 0.9 peak
    normalisation of reference data.py:162-164): 10 log10(|s|^2 / |s - a*est|^2)."""
    ref = np.asarray(ref, dtype=np.float64)
    est = np.asarray(est, dtype=np.float64)
    n = min(len(ref), len(est))
    ref, est = ref[:n], est[:n]
    a = float(np.dot(ref, est) / max(np.dot(est, est), 1e-30))
    err = ref - a * est
    return 10.0 * np.log10(max(np.dot(ref, ref), 1e-30) / max(np.dot(err, err), 1e-30))
