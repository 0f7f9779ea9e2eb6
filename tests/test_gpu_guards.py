"""GPU: no kernel writes outside the buffers it was given.  compute-sanitizer is closed on this GPU pool
(profiles/r02_sanitizer_closed.txt), so out-of-bounds writes are hunted the manual way: every output and workspace of
the C ABI is carved out of a larger allocation whose guard bands are filled with a byte pattern and checked after the
calls.  (Races are covered by the bit-reproducibility tests: graph replays, train steps and the iSTFT are compared
bit for bit across runs.)"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from svs_unet_pytorch_b200 import _lib, model as svs_model, spectral, synth, training  # noqa: E402

GUARD = 1 << 20
PATTERN = 0x5A


class Guarded:
    """`nbytes` of device memory (1024-byte aligned) with GUARD bytes of pattern on both sides."""

    def __init__(self, nbytes, dtype=torch.uint8):
        self.raw = torch.full((nbytes + 2 * GUARD + 2048,), PATTERN, dtype=torch.uint8, device="cuda")
        off = GUARD + ((-(self.raw.data_ptr() + GUARD)) % 1024)
        self.off, self.nbytes = off, nbytes
        self.view = self.raw[off:off + nbytes]
        self.t = self.view.view(dtype) if dtype != torch.uint8 else self.view

    def intact(self):
        lo, hi = self.raw[:self.off], self.raw[self.off + self.nbytes:]
        return bool((lo == PATTERN).all()) and bool((hi == PATTERN).all())


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
@pytest.mark.parametrize("batch", [3, 8])
def test_unet_forward_stays_inside_its_workspace_and_output(precision, batch):
    torch.manual_seed(0)
    net = svs_model.UNet(precision=precision).eval().cuda()
    plan = net.plan()
    lib = _lib.load()
    ws = Guarded(lib.svs_unet_workspace_bytes(plan.handle, batch))
    out = Guarded(batch * 512 * 128 * 4, torch.float32)
    x = torch.rand(batch, 1, 512, 128, device="cuda")
    iv = _lib.PatchView(x.data_ptr(), None, 512 * 128, 128, 1)
    ov = _lib.PatchView(out.t.data_ptr(), None, 512 * 128, 128, 1)
    _lib.check(lib.svs_unet_forward(plan.handle, ctypes.byref(iv), ctypes.byref(ov), None, batch, _lib.FLAG_APPLY_MASK,
                                    ws.view.data_ptr(), ws.nbytes, _lib.stream_ptr(x.device)), "svs_unet_forward")
    torch.cuda.synchronize()
    assert ws.intact() and out.intact()
    ref = net.separate(x)
    assert torch.equal(out.t.view(batch, 1, 512, 128), ref)


def test_spectral_kernels_stay_inside_their_outputs():
    songs = [synth.synth_song(7.0, seed=1)[0], np.zeros(700, dtype=np.float32), synth.synth_song(3.3, seed=2)[0]]
    batch = spectral.SongBatch.from_audio(songs)
    lib = _lib.load()
    f = batch.total_frames
    mag, phase = Guarded(f * 513 * 4, torch.float32), Guarded(f * 513 * 8, torch.float32)
    smax = Guarded(batch.n_songs * 4, torch.float32)
    st = _lib.stream_ptr(batch.audio.device)
    _lib.check(lib.svs_stft_mag_phase(batch.audio.data_ptr(), batch.sample_off.data_ptr(), batch.frame_off.data_ptr(),
                                      batch.n_songs, batch.max_frames, mag.t.data_ptr(), phase.t.data_ptr(),
                                      smax.t.data_ptr(), st), "stft")
    wave, peak = Guarded(max(batch.total_wave, 1) * 4, torch.float32), Guarded(batch.n_songs * 4, torch.float32)
    _lib.check(lib.svs_istft_ola(mag.t.data_ptr(), phase.t.data_ptr(), batch.frame_off.data_ptr(),
                                 batch.wave_off.data_ptr(), batch.n_songs, batch.max_frames, wave.t.data_ptr(),
                                 peak.t.data_ptr(), st), "istft")
    pcm = Guarded(max(batch.total_wave, 1) * 2, torch.int16)
    _lib.check(lib.svs_wave_peak_normalize_pcm16(wave.t.data_ptr(), batch.wave_off.data_ptr(), peak.t.data_ptr(),
                                                 batch.n_songs, batch.total_wave, 0.9, pcm.t.data_ptr(), st), "pcm16")
    torch.cuda.synchronize()
    for g in (mag, phase, smax, wave, peak, pcm):
        assert g.intact()
    m2, p2, _ = batch.stft()
    assert torch.equal(mag.t.view(f, 513), m2) and torch.equal(phase.t.view(f, 513, 2), p2)


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_training_step_stays_inside_its_workspace(precision):
    torch.manual_seed(0)
    net = svs_model.UNet().train().cuda()
    net.train_precision = precision
    batch = 3
    ws = Guarded(_lib.load().svs_unet_train_workspace_bytes(batch))
    net.__dict__["_train_ws"] = {batch: ws.view}                     # the step uses this workspace
    mix = torch.rand(batch, 1, 512, 128, device="cuda")
    voc = mix * torch.rand_like(mix)
    loss = training.train_step(net, mix, voc, step=False, use_graph=False)
    torch.cuda.synchronize()
    assert ws.intact() and torch.isfinite(loss).all()
