"""GPU: the three drop-in command lines end to end on a tiny synthetic song folder, checked against the
oracle's restatement of reference data.py / inference.py, plus one epoch of the train.py drop-in."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import stft_oracle as so, unet_oracle  # noqa: E402
from svs_unet_pytorch_b200 import audio_io, data, inference, model as svs_model, separate, synth, train  # noqa: E402


def _make_songs(root, n=2, seconds=13.0):
    stems = []
    for i in range(n):
        mix, voc, _ = synth.synth_song(seconds + i, seed=100 + i)
        d = os.path.join(root, f"song{i}")
        os.makedirs(d)
        audio_io.write_wav_pcm16(os.path.join(d, "mixture.wav"), mix, 8192)
        audio_io.write_wav_pcm16(os.path.join(d, "vocals.wav"), voc, 8192)
        stems.append((audio_io.load(os.path.join(d, "mixture.wav"), 8192),
                      audio_io.load(os.path.join(d, "vocals.wav"), 8192)))
    return stems


def test_to_spec_inference_to_wave(tmp_path):
    src, spec_dir, pred_dir, wav_dir = (str(tmp_path / n) for n in ("songs", "spec", "pred", "wav"))
    os.makedirs(src)
    stems = _make_songs(src)
    data.main(["--src", src, "--tar", spec_dir, "--direction", "to_spec"])
    torch.manual_seed(0)
    net = svs_model.UNet()
    ckpt = str(tmp_path / "svs_test.pth")
    torch.save({"epoch": 1, "model_state_dict": net.state_dict(), "optim": net.optim.state_dict(),
                "scheduler": None}, ckpt)                             # reference train.py:369-374 layout
    os.environ["SVS_B200_PRECISION"] = "tf32"
    try:
        inference.main(["--model_path", ckpt, "--mixture_folder", os.path.join(spec_dir, "mixture"), "--tar", pred_dir,
                        "--vocal_solo", "1"])
    finally:
        del os.environ["SVS_B200_PRECISION"]
    data.main(["--src", pred_dir, "--phase", spec_dir, "--tar", wav_dir, "--direction", "to_wave"])
    sd = net.state_dict()
    for i, (mix, voc) in enumerate(stems):
        base = f"{i:04d}_song{i}"
        spec = np.load(os.path.join(spec_dir, "mixture", base + "_spec.npy"))
        phase = np.load(os.path.join(spec_dir, "mixture", base + "_phase.npy"))
        vspec = np.load(os.path.join(spec_dir, "vocal", base + "_spec.npy"))
        rspec, rphase, _ = so.to_spec(mix)
        rvspec, _, _ = so.to_spec(mix, voc)
        assert spec.shape == rspec.shape and spec.dtype == np.float32 and spec.flags["F_CONTIGUOUS"]
        assert phase.dtype == np.complex64 and phase.shape == rphase.shape
        assert np.abs(spec - rspec).max() <= 1e-4 and np.abs(vspec - rvspec).max() <= 1e-4
        assert spec.max() == np.float32(1.0)
        pred = np.load(os.path.join(pred_dir, base + "_spec.npy"))
        rpred = unet_oracle.separate_spectrogram(sd, rspec, vocal_solo=True)
        assert pred.shape == rpred.shape and pred.dtype == np.float32 and np.all(pred[0] == 0)
        assert np.abs(pred - rpred).max() <= 1e-3
        y, sr = audio_io.read_wav(os.path.join(wav_dir, base + ".wav"))
        ry = so.to_wave(rpred, rphase)
        assert sr == 8192 and y.shape == ry.shape
        assert abs(synth.sdr_db(voc, y) - synth.sdr_db(voc, ry)) <= 0.05


def test_fused_wav_to_wav_matches_the_three_stage_flow(tmp_path):
    # one process, spectrograms stay in HBM; --emit_npy leaves the same files behind
    src, spec_dir, pred_dir, wav_dir, out_dir, npy_dir = (str(tmp_path / n) for n in
                                                          ("songs", "spec", "pred", "wav", "fused", "fused_npy"))
    os.makedirs(src)
    stems = _make_songs(src, n=3, seconds=11.0)
    torch.manual_seed(0)
    net = svs_model.UNet()
    ckpt = str(tmp_path / "svs_test.pth")
    torch.save({"epoch": 1, "model_state_dict": net.state_dict(), "optim": net.optim.state_dict(), "scheduler": None}, ckpt)
    os.environ["SVS_B200_PRECISION"] = "tf32"
    try:
        data.main(["--src", src, "--tar", spec_dir, "--direction", "to_spec"])
        inference.main(["--model_path", ckpt, "--mixture_folder", os.path.join(spec_dir, "mixture"), "--tar", pred_dir])
        data.main(["--src", pred_dir, "--phase", spec_dir, "--tar", wav_dir, "--direction", "to_wave"])
        separate.main(["--model_path", ckpt, "--src", src, "--tar", out_dir, "--emit_npy", npy_dir, "--songs_per_batch", "2"])
    finally:
        del os.environ["SVS_B200_PRECISION"]
    sd = net.state_dict()
    for i, (mix, voc) in enumerate(stems):
        base = f"{i:04d}_song{i}"
        y3, sr3 = audio_io.read_wav(os.path.join(wav_dir, base + ".wav"))
        y1, sr1 = audio_io.read_wav(os.path.join(out_dir, base + ".wav"))
        assert sr1 == sr3 == 8192 and y1.shape == y3.shape
        assert np.abs(y1 - y3).max() <= 2e-3                           # both PCM_16; conv1 path may differ (dense vs view)
        rspec, rphase, _ = so.to_spec(mix)
        ry = so.to_wave(unet_oracle.separate_spectrogram(sd, rspec, vocal_solo=True), rphase)
        assert abs(synth.sdr_db(voc, y1) - synth.sdr_db(voc, ry)) <= 0.05
        for sub, name in (("mixture", base + "_spec.npy"), ("mixture", base + "_phase.npy")):
            a, b = np.load(os.path.join(npy_dir, sub, name)), np.load(os.path.join(spec_dir, sub, name))
            assert a.dtype == b.dtype and a.shape == b.shape and a.flags["F_CONTIGUOUS"] == b.flags["F_CONTIGUOUS"]
            assert np.array_equal(a, b)
        p1, p3 = np.load(os.path.join(npy_dir, base + "_spec.npy")), np.load(os.path.join(pred_dir, base + "_spec.npy"))
        assert p1.shape == p3.shape and p1.dtype == p3.dtype and np.abs(p1 - p3).max() <= 1e-3


def test_train_cli_one_epoch(tmp_path, monkeypatch):
    src, spec_dir = str(tmp_path / "songs"), str(tmp_path / "spec")
    os.makedirs(src)
    _make_songs(src, n=2)
    data.main(["--src", src, "--tar", spec_dir, "--direction", "to_spec"])
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(train, "SAMPLES_PER_SONG", 4)
    ds = train.SpectrogramDataset(spec_dir, samples_per_song=4, device="cuda", seed=0)
    assert len(ds) == 8
    mix, voc = next(ds.batches(4))
    assert mix.shape == voc.shape == (4, 1, 512, 128) and mix.is_cuda
    train.main(["--train_folder", spec_dir, "--valid_folder", spec_dir, "--label", "t", "--epoch", "2",
                "--batch_size", "4", "--val_interval", "1", "--load_path", "none.pth"])
    ckpt = torch.load(str(tmp_path / "CKPT" / "svs_t.pth"), map_location="cpu")
    assert ckpt["epoch"] == 2 and set(ckpt) >= {"model_state_dict", "optim", "scheduler", "loss_list_total"}
    assert len(ckpt["model_state_dict"]) == 79
    assert os.path.exists(str(tmp_path / "CKPT" / "svs_best_t.pth"))
    lines = open(str(tmp_path / "LOG" / "log_t.txt")).read().split("\n")
    assert any(l.startswith("Val ") for l in lines) and float(lines[0]) > 0      # loss_plot.py format
    # resume from the checkpoint (train.py:216-237), this time on the L1-only fused CUDA-graph step
    train.main(["--train_folder", spec_dir, "--valid_folder", "missing", "--label", "t", "--epoch", "3",
                "--batch_size", "4", "--load_path", str(tmp_path / "CKPT" / "svs_t.pth"), "--mr_stft", "0"])
    assert torch.load(str(tmp_path / "CKPT" / "svs_t.pth"), map_location="cpu")["epoch"] == 3


def test_training_crops_match_the_reference_getitem(tmp_path):
    # SpectrogramDataset.crop_batch (two svs_patches_gather launches per batch) against a numpy restatement of the
    # reference's __getitem__ (train.py:99-131: DC row dropped, random shared start, zero padding when short)
    import random
    spec_dir = tmp_path / "spec"
    rng = np.random.default_rng(0)
    lens = [300, 128, 57, 129]
    for sub in ("mixture", "vocal"):
        os.makedirs(spec_dir / sub)
    specs = {}
    for i, t in enumerate(lens):
        for sub in ("mixture", "vocal"):
            a = np.asfortranarray(rng.random((513, t), dtype=np.float32))          # data.py writes F-ordered (513, T)
            np.save(spec_dir / sub / f"{i:04d}_song_spec.npy", a)
            specs[(sub, i)] = a
    ds = train.SpectrogramDataset(str(spec_dir), samples_per_song=3, device="cuda", seed=5)
    ref_rng = random.Random(5)
    idx = [0, 5, 2, 7, 3, 9, 10, 1]
    mix, voc = ds.crop_batch(idx)
    assert mix.shape == voc.shape == (8, 1, 512, 128)
    for k, i in enumerate(idx):
        s = i % 4
        m, v = specs[("mixture", s)][1:, :], specs[("vocal", s)][1:, :]
        cur = m.shape[1]
        if cur > 128:
            start = ref_rng.randint(0, cur - 128)
            m, v = m[:, start:start + 128], v[:, start:start + 128]
        else:
            m, v = np.pad(m, ((0, 0), (0, 128 - cur))), np.pad(v, ((0, 0), (0, 128 - cur)))
        assert np.array_equal(mix[k, 0].cpu().numpy(), m) and np.array_equal(voc[k, 0].cpu().numpy(), v), k
