"""CPU: host-side logic — WAV I/O, CLI surfaces, patch tables, song sharding, gloo multi-process paths."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_wav_round_trip(tmp_path):
    from svs_unet_pytorch_b200 import audio_io
    y = (0.5 * np.sin(2 * np.pi * 220 * np.arange(8192) / 8192)).astype(np.float32)
    p = str(tmp_path / "a.wav")
    audio_io.write_wav_pcm16(p, y, 8192)
    x, sr = audio_io.read_wav(p)
    assert sr == 8192 and x.shape == y.shape and np.abs(x - y).max() <= 1.0 / 32768 + 1e-7
    assert np.array_equal(audio_io.load(p, sr=8192), x)


def test_cli_flags_match_reference():
    from svs_unet_pytorch_b200 import data, inference, train
    d = data.build_parser().parse_args(["--src", "a", "--tar", "b"])
    assert (d.win_size, d.hop_size, d.sr, d.direction, d.phase) == (1024, 768, 8192, "to_spec", "-1")   # data.py:20-28
    i = inference.build_parser().parse_args(["--model_path", "m", "--tar", "t", "--mixture_folder", "f"])
    assert i.vocal_solo == 1                                                                          # inference.py:29-34
    t = train.build_parser().parse_args(["--label", "x"])
    assert (t.epoch, t.batch_size, t.val_interval, t.load_path) == (2, 2, 20, "result.pth")           # train.py:157-167


def test_fused_cli_song_discovery_and_no_cpu_path(tmp_path):
    # scripts/separate.py: same enumeration order as reference data.py:56 (sorted folder names), plain wav fallback,
    # and -- like every product entry point -- it refuses to run without a GPU instead of falling back
    from svs_unet_pytorch_b200 import _lib, audio_io, separate
    y = np.zeros(8192, dtype=np.float32)
    for name in ("b_song", "a_song", "no_mixture"):
        os.makedirs(tmp_path / "songs" / name)
    audio_io.write_wav_pcm16(str(tmp_path / "songs" / "b_song" / "mixture.wav"), y, 8192)
    audio_io.write_wav_pcm16(str(tmp_path / "songs" / "a_song" / "mixture.wav"), y, 8192)
    found = separate.find_songs(str(tmp_path / "songs"))
    assert [n for n, _ in found] == ["a_song", "b_song"]
    os.makedirs(tmp_path / "flat")
    audio_io.write_wav_pcm16(str(tmp_path / "flat" / "x.wav"), y, 8192)
    assert separate.find_songs(str(tmp_path / "flat")) == [("x", str(tmp_path / "flat" / "x.wav"))]
    a = separate.build_parser().parse_args(["--model_path", "m", "--src", "s", "--tar", "t"])
    assert (a.vocal_solo, a.sr, a.emit_npy) == (1, 8192, "")
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SvsError):
            separate.main(["--model_path", "m", "--src", str(tmp_path / "songs"), "--tar", str(tmp_path / "out")])


def test_patch_table_edge_cases():
    from svs_unet_pytorch_b200 import pipeline
    offs, valid, song = pipeline.patch_table([1, 127, 128, 129, 256], np.array([0, 1, 128, 256, 385, 641]))
    assert [int((song == s).sum()) for s in range(5)] == [1, 1, 1, 2, 2]
    assert list(valid) == [1, 127, 128, 128, 1, 128, 128]
    assert all((o - 1) % 513 == 0 for o in offs)


def test_geometry_other_than_1209_is_rejected():
    from svs_unet_pytorch_b200 import spectral
    with pytest.raises(RuntimeError, match="n_fft=1024"):
        spectral._check_geometry(1024, 256)                  # the 44.1 kHz sets of reference config.py:18-44


def test_song_sharding_is_balanced():
    from svs_unet_pytorch_b200 import sharding
    for world in (1, 2, 4, 8):
        parts = [sharding.shard_songs(150, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(150))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1      # 150 songs: 19 x 6 + 18 x 2 at 8 GPUs


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SVS_ROOT"])
from svs_unet_pytorch_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = sharding.shard_songs(11, rank, world)
# max-over-ranks timing reduction and work gather used by bench.py / the corpus driver
t = sharding.max_over_ranks(float(rank + 1), backend_device="cpu")
n = sharding.sum_over_ranks(float(len(mine)), backend_device="cpu")
# data-parallel gradient averaging on a flat buffer (what training.train_step does with NCCL)
flat = torch.full((1000,), float(rank))
sharding.average_flat_(flat)
assert abs(t - world) < 1e-9 and abs(n - 11) < 1e-9, (t, n)
assert torch.allclose(flat, torch.full((1000,), (world - 1) / 2.0))
dist.destroy_process_group()
print("ok", rank)
"""


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, SVS_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_data_parallel_epoch_order_is_a_partition_with_equal_steps():
    # train.SpectrogramDataset.epoch_order: every rank shuffles with the shared epoch seed, the order is padded to
    # a multiple of world * batch so that all ranks run the same number of full steps (else the all-reduce hangs)
    import random
    from svs_unet_pytorch_b200 import train as svs_train
    ds = svs_train.SpectrogramDataset.__new__(svs_train.SpectrogramDataset)
    ds.file_names = [f"{i:04d}" for i in range(7)]
    ds.samples_per_song = 9                                         # 63 items
    for world in (2, 3, 5, 8):
        for bs in (2, 4):
            ds.rng = random.Random(1)
            orders = [ds.epoch_order(bs, True, r, world, epoch_seed=77) for r in range(world)]
            assert len({len(o) for o in orders}) == 1 and len(orders[0]) % bs == 0
            assert len(orders[0]) // bs == ds.n_batches(bs, world)
            seen = sum(orders, [])
            assert set(seen) == set(range(63))                      # nobody is skipped
            assert len(seen) - 63 < world * bs                      # only the wrap-around padding repeats
    ds.rng = random.Random(1)
    single = ds.epoch_order(4, True, 0, 1)
    assert sorted(single) == list(range(63)) and ds.n_batches(4, 1) == 16   # reference DataLoader: last batch kept
