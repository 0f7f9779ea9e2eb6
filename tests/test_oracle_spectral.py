"""CPU: the librosa-0.10.1 restatement (oracle/stft_oracle.py) against torch.stft / torch.istft in
float64 (an independent statement of the same transforms) and the structural facts of SURVEY.md section 4.
PARITY UNPINNED at the librosa boundary: the reference ships no vectors for this half."""
import numpy as np
import torch

from oracle import stft_oracle as so
from svs_unet_pytorch_b200 import synth


def _torch_stft64(y):
    w = torch.hann_window(1024, periodic=True, dtype=torch.float64)
    return torch.stft(torch.from_numpy(y).double(), n_fft=1024, hop_length=768, window=w, center=True,
                      pad_mode="constant", return_complex=True).numpy()


def test_known_shapes():
    assert so.n_frames(245760) == 321 and so.n_frames(1474560) == 1921
    y = np.zeros(245760, dtype=np.float32)
    d = so.stft(y)
    assert d.shape == (513, 321) and d.dtype == np.complex64 and d.flags["F_CONTIGUOUS"]


def test_stft_matches_torch_float64():
    mix, _, _ = synth.synth_song(5.0, seed=1234)
    d = so.stft(mix)
    ref = _torch_stft64(mix)
    assert d.shape == ref.shape
    assert np.abs(d - ref.astype(np.complex64)).max() <= 1e-6 * np.abs(ref).max()


def test_magphase_zero_rule():
    d = np.array([[0 + 0j, 3 + 4j], [1e-30 + 0j, -2j]], dtype=np.complex64)
    mag, ph = so.magphase(d)
    assert mag.dtype == np.float32 and ph.dtype == np.complex64
    assert ph[0, 0] == 1 + 0j and np.allclose(ph[0, 1], 0.6 + 0.8j) and np.allclose(ph[1, 1], -1j)
    assert np.allclose(np.abs(ph), 1.0)


def test_to_spec_normalises_by_mixture_max():
    mix, voc, _ = synth.synth_song(3.0, seed=7)
    spec_m, ph_m, norm = so.to_spec(mix)
    spec_v, _, norm_v = so.to_spec(mix, voc)
    assert spec_m.max() == np.float32(1.0) and norm == norm_v
    assert spec_v.max() < 1.0 + 1e-6
    z, _, n0 = so.to_spec(np.zeros(4096, dtype=np.float32))
    assert n0 == 1 and np.all(z == 0)


def test_istft_matches_torch_and_round_trips():
    rng = np.random.default_rng(0)
    y = (0.1 * rng.standard_normal(768 * 40)).astype(np.float32)
    d = so.stft(y)
    yr = so.istft(d)
    assert yr.dtype == np.float32 and yr.shape[0] == 768 * (d.shape[1] - 1) == len(y)
    assert np.abs(yr - y).max() < 1e-6
    w = torch.hann_window(1024, periodic=True, dtype=torch.float64)
    ref = torch.istft(torch.from_numpy(d.astype(np.complex128)), n_fft=1024, hop_length=768, window=w,
                      center=True, length=len(y)).numpy()
    assert np.abs(yr - ref).max() < 1e-6


def test_envelope_never_zero_after_trim():
    env = so.window_sumsquare(10)[512:512 + 768 * 9]
    assert env.min() > 0.04 and env.max() <= 1.0 + 1e-6


def test_to_wave_peak_normalises():
    mix, _, _ = synth.synth_song(3.0, seed=5)
    spec, ph, _ = so.to_spec(mix)
    y = so.to_wave(spec, ph)
    assert abs(np.abs(y).max() - 0.9) < 1e-6
    assert synth.sdr_db(mix[: len(y)], y) > 60.0


# ---------------------------------------------------------------------------------------------
# Analytic known-answer vectors: a third, library-independent check of the restatement (the librosa boundary
# itself stays UNPINNED: librosa is absent here and the reference ships no vectors).
# Periodic Hann has the 3-term spectrum W[0] = N/2, W[+-1] = -N/4, so for x[n] = cos(2 pi k0 n / N):
#   X_t[k0] = N/4 * exp(2 pi i k0 s_t / N),  X_t[k0 +- 1] = -N/8 * (same phasor),  all other bins 0,
# for every frame whose window lies inside the signal; s_t = 768 t - 512 is the frame's first sample.

def sinusoid_kat(k0: int, n_frames: int = 12):
    n = 768 * (n_frames - 1)
    y = np.cos(2 * np.pi * k0 * np.arange(n) / 1024.0).astype(np.float32)
    t = np.arange(1, n_frames - 1)                                    # interior frames (no centre padding)
    phasor = np.exp(2j * np.pi * k0 * (768 * t - 512) / 1024.0)
    expect = np.zeros((513, len(t)), dtype=np.complex128)
    expect[k0] = 256.0 * phasor
    expect[k0 - 1] = -128.0 * phasor
    expect[k0 + 1] = -128.0 * phasor
    return y, t, expect


def impulse_kat(n0: int = 3000, n_frames: int = 8):
    n = 768 * (n_frames - 1)
    y = np.zeros(n, dtype=np.float32)
    y[n0] = 1.0
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1024) / 1024.0)
    expect = np.zeros((513, n_frames), dtype=np.complex128)
    k = np.arange(513)
    for t in range(n_frames):
        m = n0 - (768 * t - 512)                                      # position of the impulse inside frame t
        if 0 <= m < 1024:
            expect[:, t] = w[m] * np.exp(-2j * np.pi * k * m / 1024.0)
    return y, expect


def test_stft_bin_centred_sinusoids_known_answer():
    for k0 in (2, 37, 256, 510):
        y, t, expect = sinusoid_kat(k0)
        d = so.stft(y)[:, t]
        assert np.abs(d - expect).max() <= 2e-5 * 256.0, k0          # float32 input quantisation only
        mag, ph = so.magphase(so.stft(y))
        assert np.abs(mag[k0, t] - 256.0).max() < 1e-2 and np.abs(mag[k0 + 1, t] - 128.0).max() < 1e-2
        assert np.abs(ph[k0, t] - expect[k0] / 256.0).max() < 1e-4


def test_stft_impulse_known_answer_and_round_trip():
    y, expect = impulse_kat()
    d = so.stft(y)
    assert np.abs(d - expect).max() <= 1e-6
    yr = so.istft(d)
    assert np.abs(yr - y).max() < 1e-6                                # sum_t w^2 / envelope == 1 exactly where covered
