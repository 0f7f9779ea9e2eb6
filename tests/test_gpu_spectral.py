"""GPU: STFT / magphase / iSTFT kernels against the CPU oracle (oracle/stft_oracle.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import stft_oracle as so  # noqa: E402
from svs_unet_pytorch_b200 import spectral, synth  # noqa: E402


def _songs():
    rng = np.random.default_rng(0)
    out = [synth.synth_song(30.0, seed=1234)[0],                       # config 1: 245,760 samples -> 321 frames
           (0.1 * rng.standard_normal(768 * 7 + 5)).astype(np.float32),   # ragged, len % 768 != 0
           np.zeros(2000, dtype=np.float32),                            # silent song: norm == 0 -> 1, phase 1+0j
           (0.5 * rng.standard_normal(700)).astype(np.float32)]        # shorter than one hop: a single frame
    return out


def test_stft_mag_phase_matches_oracle():
    songs = _songs()
    batch = spectral.SongBatch.from_audio(songs)
    mag, phase, smax = batch.stft()
    torch.cuda.synchronize()
    assert batch.frames == [so.n_frames(len(s)) for s in songs]
    for i, y in enumerate(songs):
        d = so.stft(y)
        rmag, rph = so.magphase(d)
        got = batch.song_spec(mag, i).cpu().numpy()
        assert got.shape == rmag.shape
        scale = max(float(rmag.max()), 1e-30)
        # north_star tolerance: STFT magnitude within 1e-4 relative (to the spectrum max, SURVEY section 4 caution i)
        assert np.abs(got - rmag).max() <= 1e-4 * scale, (i, np.abs(got - rmag).max(), scale)
        assert abs(float(smax[i]) - float(rmag.max())) <= 1e-4 * scale
        a, b = int(batch.frame_off_host[i]), int(batch.frame_off_host[i + 1])
        ph = torch.view_as_complex(phase[a:b]).cpu().numpy().T
        assert np.allclose(np.abs(ph), 1.0, atol=1e-5)                # unit phasors (1+0j where mag == 0)
        strong = rmag > 1e-3 * scale                                  # phase is ill-conditioned in empty bins
        if strong.any():
            assert np.abs(ph[strong] - rph[strong]).max() < 2e-3
    z = batch.song_spec(mag, 2).cpu().numpy()
    assert np.all(z == 0) and float(smax[2]) == 0.0
    a = int(batch.frame_off_host[2])
    assert torch.all(phase[a, :, 0] == 1.0) and torch.all(phase[a, :, 1] == 0.0)


def test_librosa_shaped_calls():
    y = synth.synth_song(5.0, seed=3)[0]
    d = spectral.stft(y, n_fft=1024, hop_length=768)
    ref = so.stft(y)
    assert d.shape == ref.shape == (513, 54) and d.dtype == np.complex64 and d.flags["F_CONTIGUOUS"]
    assert np.abs(d - ref).max() <= 1e-4 * np.abs(ref).max()
    mag, ph = spectral.magphase(d)
    rmag, rph = so.magphase(d)
    assert mag.dtype == np.float32 and ph.dtype == np.complex64
    assert np.abs(mag - rmag).max() <= 1e-6 * rmag.max() and np.abs(ph - rph).max() < 1e-5
    yr = spectral.istft(ref, win_length=1024, hop_length=768)
    ryr = so.istft(ref)
    assert yr.shape == ryr.shape and yr.dtype == np.float32
    assert np.abs(yr - ryr).max() < 1e-5
    with pytest.raises(RuntimeError):
        spectral.stft(y, n_fft=1024, hop_length=256)                  # other config.py sets are rejected


def test_normalize_and_round_trip_batch():
    songs = _songs()
    batch = spectral.SongBatch.from_audio(songs)
    mag, phase, smax = batch.stft()
    raw = mag.clone()
    batch.normalize(mag, smax)
    for i in range(len(songs)):
        n = float(smax[i]) or 1.0
        got = batch.song_spec(mag, i).cpu().numpy()
        exp = batch.song_spec(raw, i).cpu().numpy() / np.float32(n)    # data.py:105
        assert np.array_equal(got, exp)
    wave, peak = batch.istft(raw, phase)
    wave_n = wave.clone()
    from svs_unet_pytorch_b200 import _lib
    _lib.wave_peak_normalize_raw(wave_n, batch.wave_off, peak, batch.n_songs, 0.9)
    torch.cuda.synchronize()
    for i, y in enumerate(songs):
        got = batch.song_wave(wave, i).cpu().numpy()
        ref = so.istft(so.stft(y))
        assert got.shape == ref.shape
        if len(ref):
            assert np.abs(got - ref).max() < 2e-6 * max(1.0, np.abs(ref).max())
            assert np.abs(got - y[: len(got)]).max() < 1e-5             # STFT -> iSTFT round trip
            assert abs(float(peak[i]) - np.abs(ref).max()) < 1e-5
            gn = batch.song_wave(wave_n, i).cpu().numpy()
            if np.abs(ref).max() > 0:
                assert abs(np.abs(gn).max() - 0.9) < 1e-5               # data.py:163-164


def test_istft_is_bit_reproducible():
    songs = [synth.synth_song(180.0, seed=1234 + i)[0] for i in range(2)]   # config 3 size: 1921 frames
    batch = spectral.SongBatch.from_audio(songs)
    mag, phase, _ = batch.stft()
    w1, _ = batch.istft(mag, phase)
    w2, _ = batch.istft(mag, phase)
    assert torch.equal(w1, w2)
    assert batch.frames == [1921, 1921] and w1.numel() == 2 * 1474560
    for i, y in enumerate(songs):
        got = batch.song_wave(w1, i).cpu().numpy()
        assert np.abs(got - y).max() < 1e-5


def test_cuda_stft_matches_analytic_known_answers():
    # library-independent vectors (tests/test_oracle_spectral.py): bin-centred sinusoids and an impulse
    from test_oracle_spectral import impulse_kat, sinusoid_kat
    for k0 in (2, 37, 256, 510):
        y, t, expect = sinusoid_kat(k0)
        d = spectral.stft(y)
        assert np.abs(d[:, t] - expect).max() <= 1e-4 * 256.0, k0    # north_star: 1e-4 of the spectrum max
    y, expect = impulse_kat()
    d = spectral.stft(y)
    assert np.abs(d - expect).max() <= 1e-4
    yr = spectral.istft(d.astype(np.complex64))
    assert np.abs(yr - y).max() < 1e-5


def test_pcm16_boundary_matches_the_float_path():
    # reference data.py:78 (librosa.load of a PCM_16 file = int16 / 32768) and data.py:166 (sf.write PCM_16 =
    # lrintf(y * 0x7FFF)): the int16 entry points must equal "convert on the host, then the float kernels"
    rng = np.random.default_rng(5)
    songs16 = [rng.integers(-20000, 20000, size=n, dtype=np.int16) for n in (768 * 9 + 3, 5000, 768 * 40)]
    songs32 = [s.astype(np.float32) / 32768.0 for s in songs16]
    b32 = spectral.SongBatch.from_audio(songs32)
    pcm = torch.from_numpy(np.concatenate(songs16)).cuda()
    b16 = spectral.SongBatch(pcm, [len(s) for s in songs16])
    m32, p32, x32 = b32.stft()
    m16, p16, x16 = b16.stft()
    assert torch.equal(m32, m16) and torch.equal(p32, p16) and torch.equal(x32, x16)    # same arithmetic after the load
    wave, peak = b32.istft(m32, p32, peak_normalize=False)
    q, _ = b32.istft(m32, p32, peak_normalize=True, pcm16=True)
    assert q.dtype == torch.int16 and q.shape == wave.shape
    norm, _ = b32.istft(m32, p32, peak_normalize=True)
    ref = np.clip(np.rint(norm.cpu().numpy().astype(np.float64) * 32767.0), -32768, 32767)
    assert np.abs(q.cpu().numpy().astype(np.int64) - ref.astype(np.int64)).max() <= 1   # fp32 vs fp64 product at ties
    assert int(q.abs().max()) == int(round(0.9 * 32767))


def test_gpu_resampler_matches_scipy_polyphase_oracle():
    # 44.1 kHz stereo PCM_16 -> mono 8192 Hz (data.py:78) on the GPU against
    # scipy.signal.resample_poly in float64 after the same / 32768 and channel mean; ragged two-song batch
    from oracle import resample_oracle
    from svs_unet_pytorch_b200 import resample
    rng = np.random.default_rng(7)
    t = np.arange(44100 * 3) / 44100.0
    s1 = np.stack([0.4 * np.sin(2 * np.pi * 440 * t) + 0.05 * rng.standard_normal(len(t)),
                   0.3 * np.sin(2 * np.pi * 1000 * t) + 0.05 * rng.standard_normal(len(t))], axis=1)
    s2 = 0.3 * rng.standard_normal((50001, 2))
    pcm = [np.clip(np.rint(s * 32767), -32768, 32767).astype(np.int16) for s in (s1, s2)]
    dev_pcm = torch.from_numpy(np.concatenate([p.reshape(-1) for p in pcm])).cuda()
    out, lens = resample.resample_songs(dev_pcm, [len(p) for p in pcm], 2, 44100, 8192)
    out = out.cpu().numpy()
    off = 0
    for p, n in zip(pcm, lens):
        ref = resample_oracle.load_like(p, 44100, 8192)
        assert n == len(ref)
        assert np.abs(out[off:off + n] - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
        off += n
    # float32 mono input, another ratio (48 kHz -> 8192 Hz), through the numpy-shaped call
    x = (0.2 * rng.standard_normal(48000)).astype(np.float32)
    y = resample.resample(x, 48000, 8192)
    ref = resample_oracle.load_like(x, 48000, 8192)
    assert y.shape == ref.shape and np.abs(y - ref).max() <= 2e-6


@pytest.mark.parametrize("n_songs,seed", [(1, 0), (7, 1), (97, 2), (300, 3)])
def test_istft_ragged_batch_equals_song_by_song(n_songs, seed):
    # The iSTFT splits the batch's frames evenly over a one-wave grid, in runs that cross song borders (istft.cu).  A
    # frame's arithmetic does not depend on where the runs are cut, so a ragged batch must reproduce, BIT FOR BIT, what
    # each song gives alone -- from one-frame songs up to batches with more songs than CTAs.
    rng = np.random.default_rng(seed)
    lengths = rng.integers(1, 768 * 40, size=n_songs)
    lengths[rng.integers(0, n_songs)] = 700                           # a single-frame song somewhere
    songs = [(0.2 * rng.standard_normal(int(n))).astype(np.float32) for n in lengths]
    batch = spectral.SongBatch.from_audio(songs)
    mag, phase, _ = batch.stft()
    wave, peak = batch.istft(mag, phase)
    torch.cuda.synchronize()
    for i in rng.permutation(n_songs)[:12]:
        one = spectral.SongBatch.from_audio([songs[i]])
        m1, p1, _ = one.stft()
        w1, k1 = one.istft(m1, p1)
        assert torch.equal(batch.song_spec(mag, i), one.song_spec(m1, 0))
        got = batch.song_wave(wave, i)
        assert got.numel() == one.wave_lengths[0] and torch.equal(got, one.song_wave(w1, 0))
        assert float(peak[i]) == float(k1[0])
