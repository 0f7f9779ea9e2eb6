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


@pytest.mark.parametrize("staged", [True, False])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 1e-2), ("tf32", 1e-3)])
def test_full_song_round_trip_matches_oracle(precision, tol, staged):
    # BASELINE configs[0] (30 s) + a ragged second song, both vocal_solo settings
    songs = [synth.synth_song(30.0, seed=1234), synth.synth_song(13.7, seed=99)]
    torch.manual_seed(0)
    net = svs_model.UNet(precision=precision).eval().cuda()
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    sep = pipeline.Separator(net, max_batch=4, staged=staged)
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


def test_patch_staging_matches_indexing():
    # svs_patches_gather / svs_patches_scatter against plain indexing (reference inference.py:74-97, 110-127)
    from svs_unet_pytorch_b200 import _lib
    frames = [321, 128, 1, 200]
    frame_off = np.concatenate([[0], np.cumsum(frames)])
    offs, valid, song = pipeline.patch_table(frames, frame_off)
    g = torch.Generator().manual_seed(3)
    spec = torch.rand(int(frame_off[-1]), 513, generator=g).cuda()
    norm = torch.tensor([2.0, 0.0, 0.5, 3.0], device="cuda")[torch.from_numpy(song).cuda().long()]   # 0 -> 1
    d_off, d_valid = torch.from_numpy(offs).cuda(), torch.from_numpy(valid).cuda()
    patches = _lib.patches_gather_raw(spec, d_off, d_valid, norm)
    assert patches.shape == (len(offs), 1, 512, 128)
    for p in range(len(offs)):
        f0 = (int(offs[p]) - 1) // 513
        nrm = float(norm[p]) or 1.0
        ref = torch.zeros(512, 128, device="cuda")
        blk = spec[f0: f0 + valid[p], 1:]
        ref[:, : valid[p]] = torch.div(blk, torch.full_like(blk, nrm)).T      # true division, as data.py:85
        assert torch.equal(patches[p, 0], ref), p
    out = torch.full_like(spec, -1.0)
    _lib.patches_scatter_raw(patches, d_off, d_valid, out, dc_zero=True)
    assert torch.all(out[:, 0] == 0)
    for p in range(len(offs)):
        f0 = (int(offs[p]) - 1) // 513
        assert torch.equal(out[f0: f0 + valid[p], 1:], patches[p, 0, :, : valid[p]].T)
    assert not torch.any(out == -1.0)                                  # every frame row written in full
    same = _lib.patches_gather_raw(spec, d_off[:1], None, None)        # NULL in_frames / norm: a full, unscaled patch
    assert torch.equal(same[0, 0], spec[:128, 1:].T)


def test_staged_and_view_paths_agree_on_a_corpus_slice():
    songs = [synth.synth_song(20.0, seed=5)[0], synth.synth_song(31.0, seed=6)[0]]
    torch.manual_seed(0)
    net = svs_model.UNet(precision="bf16").eval().cuda()
    waves = []
    for staged in (True, False):
        sep = pipeline.Separator(net, max_batch=3, staged=staged)
        waves.append(sep.separate(songs))
    for a, b in zip(*waves):
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= 2e-2 * 0.9                       # conv1 differs (tensor-core vs CUDA-core path)


def test_song_streamer_matches_separate():
    songs = [synth.synth_song(9.0 + i, seed=20 + i)[0] for i in range(5)]
    torch.manual_seed(0)
    net = svs_model.UNet(precision="bf16").eval().cuda()
    ref = pipeline.Separator(net).separate(songs)
    lengths = [len(s) for s in songs]
    host_in = torch.from_numpy(np.concatenate(songs).astype(np.float32)).pin_memory()
    host_out = torch.empty(sum(768 * (n // 768) for n in lengths), dtype=torch.float32).pin_memory()
    wl = pipeline.SongStreamer(net, songs_per_chunk=2).run(host_in, lengths, host_out)
    torch.cuda.synchronize()
    off = 0
    for r, n in zip(ref, wl):
        assert n == len(r)
        # same kernels; split-K factors depend on the UNet batch size, so only the last bits may differ
        assert np.abs(host_out[off:off + n].numpy() - r).max() <= 2e-3
        off += n


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
    # a second run on the same streamer (slot events recorded by the first), more steps than staging slots, inputs
    # in another order
    hout2 = [torch.zeros(4, 1, 512, 128).pin_memory() for _ in range(11)]
    order = [3, 1, 4, 0, 2, 2, 0, 4, 1, 3, 0]
    st.run([hin[i] for i in order], hout2)
    torch.cuda.synchronize()
    for i, b in zip(order, hout2):
        assert torch.equal(b, hout[i])
    with pytest.raises(Exception):
        st.run([torch.rand(3, 1, 512, 128)], [torch.empty(3, 1, 512, 128)])     # wrong batch: rejected, not truncated
    # the minimum staging depth (two slots): every upload waits for the forward two steps back
    st2 = pipeline.PatchStreamer(net, batch=4, vocal_solo=True, n_buf=2)
    hout3 = [torch.zeros(4, 1, 512, 128).pin_memory() for _ in range(7)]
    st2.run([hin[i % 5] for i in range(7)], hout3)
    torch.cuda.synchronize()
    for i, b in enumerate(hout3):
        assert torch.equal(b, hout[i % 5])


@pytest.mark.parametrize("precision,tol", [("bf16", 1e-2), ("tf32", 1e-3)])
def test_three_minute_round_trip_matches_oracle(precision, tol):
    # BASELINE configs[2]: one synthetic 180 s mixture (1,474,560 samples, 1,921 frames, 16 patches) through the
    # production Separator (512-patch UNet batches, staged patches) against to_spec -> inference loop -> to_wave
    mix, voc, _ = synth.synth_song(180.0, seed=1234)
    torch.manual_seed(0)
    net = svs_model.UNet(precision=precision).eval().cuda()
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    sep = pipeline.Separator(net, max_batch=512)
    batch = spectral.SongBatch.from_audio([mix])
    wave, peak, mag, phase, out_mag = sep.separate_batch(batch, vocal_solo=True, return_spec=True)
    spec_ref, pred_ref, y_ref = _oracle_song(sd, mix, True)
    got_spec = batch.song_spec(mag, 0).cpu().numpy()
    assert got_spec.shape == spec_ref.shape == (513, 1921)
    assert np.abs(got_spec - spec_ref).max() <= 1e-4                  # normalised: max == 1
    got = batch.song_spec(out_mag, 0).cpu().numpy()
    assert np.abs(got - pred_ref).max() <= tol
    y = batch.song_wave(wave, 0).cpu().numpy()
    assert y.shape == y_ref.shape == (1474560,)
    d = abs(synth.sdr_db(voc, y) - synth.sdr_db(voc, y_ref))
    assert d <= 0.05, d                                               # north_star: within 0.05 dB SDR
    assert synth.sdr_db(y_ref, y) > (30.0 if precision == "bf16" else 50.0)


def test_rank_shard_of_the_corpus_matches_per_song_and_oracle():
    # the N = 8 geometry of BASELINE configs[3]: 19 three-minute songs = 304 patches in ONE ragged 512-slot batch;
    # every song must come out as it does alone, and two of them are checked against the oracle
    songs = [synth.synth_song(180.0, seed=1234 + 8 * i) for i in range(19)]
    torch.manual_seed(0)
    net = svs_model.UNet(precision="bf16").eval().cuda()
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    sep = pipeline.Separator(net, max_batch=512)
    waves = sep.separate([s[0] for s in songs])
    assert len(waves) == 19 and all(len(w) == 1474560 for w in waves)
    for i in (0, 7, 18):
        alone = sep.separate([songs[i][0]])[0]
        # same kernels; tile / split-K choices depend on the UNet batch (304 vs 16 patches)
        assert np.abs(alone - waves[i]).max() <= 2e-3
    for i in (3, 18):
        mix, voc, _ = songs[i]
        _, _, y_ref = _oracle_song(sd, mix, True)
        assert abs(synth.sdr_db(voc, waves[i]) - synth.sdr_db(voc, y_ref)) <= 0.05


def test_song_streamer_overlapping_unet_sections_do_not_share_a_workspace():
    # several equal-size multi-patch chunks on four streams with a SLOW UNet (fp32 CUDA-core mode) so that the
    # UNet sections of different streams certainly overlap: each stream must own its activation workspace
    songs = [synth.synth_song(25.0, seed=40 + i)[0] for i in range(8)]          # 3 patches each
    torch.manual_seed(0)
    net = svs_model.UNet(precision="fp32").eval().cuda()
    ref = pipeline.Separator(net, max_batch=6).separate(songs)
    lengths = [len(s) for s in songs]
    host_in = torch.from_numpy(np.concatenate(songs).astype(np.float32)).pin_memory()
    host_out = torch.empty(sum(768 * (n // 768) for n in lengths), dtype=torch.float32).pin_memory()
    st = pipeline.SongStreamer(net, songs_per_chunk=2, n_streams=4)
    for _ in range(2):
        wl = st.run(host_in, lengths, host_out)
        torch.cuda.synchronize()
        off = 0
        for r, n in zip(ref, wl):
            assert np.abs(host_out[off:off + n].numpy() - r).max() <= 1e-5
            off += n


def test_song_streamer_pcm16_in_and_out():
    songs = [synth.synth_song(14.0 + i, seed=60 + i)[0] for i in range(4)]
    pcm = [np.clip(np.rint(s * 32768.0), -32768, 32767).astype(np.int16) for s in songs]    # what a PCM_16 .wav holds
    torch.manual_seed(0)
    net = svs_model.UNet(precision="bf16").eval().cuda()
    ref = pipeline.Separator(net).separate([p.astype(np.float32) / 32768.0 for p in pcm])  # float path, same samples
    lengths = [len(s) for s in songs]
    host_in = torch.from_numpy(np.concatenate(pcm)).pin_memory()
    host_out = torch.empty(sum(768 * (n // 768) for n in lengths), dtype=torch.int16).pin_memory()
    wl = pipeline.SongStreamer(net, songs_per_chunk=2).run(host_in, lengths, host_out)
    torch.cuda.synchronize()
    off = 0
    for r, n in zip(ref, wl):
        want = np.rint(r.astype(np.float64) * 32767.0)
        got = host_out[off:off + n].numpy().astype(np.float64)
        assert np.abs(got - want).max() <= 2e-3 * 32767                # split-K / batch-size dependent last bits
        off += n
