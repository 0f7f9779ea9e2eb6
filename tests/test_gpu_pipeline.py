"""GPU: the fused song pipeline (STFT -> UNet -> iSTFT) against the oracle's three-stage restatement of
reference data.py / inference.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import stft_oracle as so, unet_oracle  # noqa: E402
from svs_unet_pytorch_b200 import model as svs_model, pipeline, spectral, synth  # noqa: E402


def _oracle_song(sd, mix, vocal_solo=True):
    spec, phase, _ = so.to_spec(mix)
    pred = unet_oracle.separate_spectrogram(sd, spec, vocal_solo=vocal_solo)
    return spec, pred, so.to_wave(pred, phase)


def test_patch_table_matches_reference_segmentation():
    # reference inference.py:75,88: T//128 + 1 segments, the empty one skipped
    offs, valid, song = pipeline.patch_table([321, 1921, 128, 1], np.array([0, 321, 2242, 2370, 2371]))
    assert list(valid[:3]) == [128, 128, 65]                           # 30 s song: 3 patches
    assert (song == 1).sum() == 16 and valid[song == 1][-1] == 1       # 3 min song: 15 full + 1 single-frame patch
    assert (song == 2).sum() == 1 and (song == 3).sum() == 1           # T % 128 == 0: no empty patch
    assert offs[0] == 1 and offs[1] == 128 * 513 + 1 and offs[3] == 321 * 513 + 1


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 1e-2), ("tf32", 1e-3)])
def test_full_song_round_trip_matches_oracle(precision, tol):
    # BASELINE configs[0] (30 s) + a ragged second song, both vocal_solo settings
    songs = [synth.synth_song(30.0, seed=1234), synth.synth_song(13.7, seed=99)]
    torch.manual_seed(0)
    net = svs_model.UNet(precision=precision).eval().cuda()
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    sep = pipeline.Separator(net, max_batch=4)
    for vocal_solo in (True, False):
        batch = spectral.SongBatch.from_audio([s[0] for s in songs])
        wave, peak, mag, phase, out_mag = sep.separate_batch(batch, vocal_solo=vocal_solo, return_spec=True)
        for i, (mix, voc, acc) in enumerate(songs):
            spec_ref, pred_ref, y_ref = _oracle_song(sd, mix, vocal_solo)
            got = batch.song_spec(out_mag, i).cpu().numpy()
            assert got.shape == pred_ref.shape
            assert np.all(got[0] == 0)                                # DC row re-inserted as zeros
            assert np.abs(got - pred_ref).max() <= tol
            y = batch.song_wave(wave, i).cpu().numpy()
            assert y.shape == y_ref.shape
            assert abs(np.abs(y).max() - 0.9) < 1e-4                  # data.py:163-164
            target = voc if vocal_solo else acc
            d = abs(synth.sdr_db(target, y) - synth.sdr_db(target, y_ref))
            assert d <= 0.05, d                                       # north_star: within 0.05 dB SDR


def test_host_streamer_matches_direct_forward():
    torch.manual_seed(0)
    net = svs_model.UNet(precision="bf16").eval().cuda()
    st = pipeline.PatchStreamer(net, batch=4, vocal_solo=True)
    hin = [torch.rand(4, 1, 512, 128).pin_memory() for _ in range(5)]
    hout = [torch.empty(4, 1, 512, 128).pin_memory() for _ in range(5)]
    st.run(hin, hout)
    torch.cuda.synchronize()
    for a, b in zip(hin, hout):
        ref = net.separate(a.cuda(), vocal_solo=True).cpu()
        assert torch.equal(b, ref)
