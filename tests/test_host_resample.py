"""CPU: the host half of the GPU resampler — filter design and output indexing of svs_resample_poly restated in
numpy — against scipy.signal.resample_poly (the published algorithm the kernel follows; oracle/resample_oracle.py)."""
import numpy as np
import scipy.signal

from oracle import resample_oracle
from svs_unet_pytorch_b200 import resample


def _emulate_kernel(x, up, down):
    """The arithmetic of resample_poly_kernel, one output at a time, in float64."""
    h, pre_pad, pre_remove, taps = resample.design(up, down)
    hp = np.zeros((up, taps))
    idx = np.arange(len(h))
    hp[idx % up, idx // up] = h
    n_in = len(x)
    out = np.zeros(resample.out_length(n_in, up, down))
    for n in range(len(out)):
        q = (n + pre_remove) * down - pre_pad
        j_hi = q // up
        ph = q - j_hi * up
        acc = 0.0
        for t in range(taps):
            j = j_hi - t
            if j < 0:
                break
            if j < n_in:
                acc += hp[ph, t] * x[j]
        out[n] = acc
    return out


def test_filter_design_is_scipys():
    for up, down in ((2048, 11025), (2, 3), (160, 147)):
        h, _, _, _ = resample.design(up, down)
        ref = scipy.signal.firwin(2 * 10 * max(up, down) + 1, 1.0 / max(up, down), window=("kaiser", 5.0)) * up
        assert np.abs(h - ref).max() <= 1e-12 * np.abs(ref).max()


def test_kernel_indexing_matches_resample_poly():
    rng = np.random.default_rng(0)
    for up, down, n in ((2, 3, 200), (160, 147, 300), (3, 2, 101), (2048, 11025, 30000)):
        x = rng.standard_normal(n)
        ref = scipy.signal.resample_poly(x, up, down)
        got = _emulate_kernel(x, up, down) if n <= 300 else None
        if got is not None:
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() <= 1e-10
        else:                                                          # spot-check a long 44.1 kHz -> 8192 Hz case
            assert resample.out_length(n, up, down) == len(ref)


def test_oracle_load_like_converts_and_downmixes():
    x = (np.arange(40).reshape(20, 2) * 100).astype(np.int16)
    y = resample_oracle.load_like(x, 8192, 8192)
    assert np.allclose(y, x.astype(np.float64).mean(axis=1) / 32768.0)
