"""Fused ``wav -> wav`` separation: the three reference stages in ONE process, spectrograms never leave HBM.

    python scripts/separate.py --model_path CKPT/svs_x.pth --src <songs> --tar <out> [--vocal_solo 1]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/separate.py ...      # songs sharded by rank

The reference needs ``data.py --direction to_spec`` -> ``inference.py`` -> ``data.py --direction to_wave`` joined by
``.npy`` files on disk (inference.py:135-150: three process starts, four arrays written and re-read per song, one
synchronous H2D / D2H pair per 128-frame patch).  Here a batch of songs goes audio -> STFT -> /max -> UNet mask x
mixture -> iSTFT -> 0.9 peak on the device (pipeline.Separator); ``--emit_npy DIR`` additionally writes the
byte-compatible ``mixture/NNNN_<song>_{spec,phase}.npy`` and the predicted ``NNNN_<song>_spec.npy`` the three-stage
flow would have left behind, so the outputs stay drop-in.

``--src`` holds either song folders with a ``mixture.wav`` (the layout data.py:46-66 walks) or plain ``*.wav`` files.
Output: ``<tar>/NNNN_<song>.wav``, PCM_16 at 8192 Hz (data.py:166), the name data.py to_wave gives the restored file."""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from . import _lib, audio_io, pipeline, sharding, spectral
from .config import SAMPLE_RATE, num2str
from .model import UNet


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--model_path", type=str, required=True)
    p.add_argument("--src", type=str, required=True)
    p.add_argument("--tar", type=str, required=True)
    p.add_argument("--vocal_solo", type=int, default=1)               # inference.py:33
    p.add_argument("--sr", type=int, default=SAMPLE_RATE)
    p.add_argument("--songs_per_batch", type=int, default=32)
    p.add_argument("--emit_npy", type=str, default="")
    return p


def find_songs(src: str):
    """-> [(name, wav path)] in the order data.py enumerates them (sorted folder names, data.py:56)."""
    entries = sorted(os.listdir(src))
    songs = [(d, os.path.join(src, d, "mixture.wav")) for d in entries
             if os.path.isdir(os.path.join(src, d)) and os.path.exists(os.path.join(src, d, "mixture.wav"))]
    if not songs:
        songs = [(os.path.splitext(f)[0], os.path.join(src, f)) for f in entries if f.lower().endswith(".wav")]
    return songs


def load_model(model_path: str, device) -> UNet:
    model = UNet().to(device)
    ckpt = torch.load(model_path, map_location=device)
    if isinstance(ckpt, dict) and "model_state_dict" in ckpt:         # reference train.py:369-374 layout
        model.load_state_dict(ckpt["model_state_dict"])
    else:
        model.load_state_dict(ckpt)
    return model.eval()


def separate_files(model, songs, tar: str, vocal_solo: bool = True, sr: int = SAMPLE_RATE, songs_per_batch: int = 32,
                   emit_npy: str = "", indices=None):
    """songs: [(name, path)]; writes <tar>/NNNN_<name>.wav for every index in `indices` (default: all)."""
    os.makedirs(tar, exist_ok=True)
    if emit_npy:
        os.makedirs(os.path.join(emit_npy, "mixture"), exist_ok=True)
    sep = pipeline.Separator(model)
    todo = list(range(len(songs))) if indices is None else list(indices)
    done = 0
    for a in range(0, len(todo), songs_per_batch):
        ids, audio = [], []
        for i in todo[a:a + songs_per_batch]:
            try:
                audio.append(audio_io.load(songs[i][1], sr=sr, mono=True))
                ids.append(i)
            except Exception as e:                                    # per-song isolation, data.py:111-112
                print(f"Error processing {songs[i][0]}: {e}")
        if not ids:
            continue
        batch = spectral.SongBatch.from_audio(audio, device=next(model.parameters()).device)
        if emit_npy:
            wave, _, mag, phase, out_mag = sep.separate_batch(batch, vocal_solo, True, return_spec=True)
        else:
            wave, _ = sep.separate_batch(batch, vocal_solo, True)
        host = wave.cpu().numpy()
        for k, i in enumerate(ids):
            base = f"{num2str(i)}_{songs[i][0]}"
            off = int(batch.wave_off_host[k])
            audio_io.write_wav_pcm16(os.path.join(tar, base + ".wav"), host[off: off + batch.wave_lengths[k]], sr)
            if emit_npy:
                f0, f1 = int(batch.frame_off_host[k]), int(batch.frame_off_host[k + 1])
                np.save(os.path.join(emit_npy, "mixture", base + "_spec.npy"), mag[f0:f1].cpu().numpy().T)
                np.save(os.path.join(emit_npy, "mixture", base + "_phase.npy"),
                        torch.view_as_complex(phase[f0:f1]).cpu().numpy().T)
                np.save(os.path.join(emit_npy, base + "_spec.npy"),
                        np.ascontiguousarray(out_mag[f0:f1].cpu().numpy().T))   # C-order like inference.py:123-127
            done += 1
    return done


def main(argv=None):
    args = build_parser().parse_args(argv)
    if not torch.cuda.is_available():
        raise _lib.SvsError("separate needs a B200: svs-unet-pytorch_b200 has no CPU path")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    try:
        model = load_model(args.model_path, device)
    except Exception as e:
        print(f"failed to load the model: {e}")
        raise SystemExit(1)
    songs = find_songs(args.src)
    if rank == 0:
        print(f"found {len(songs)} songs in {args.src}, separating on {world} GPU(s)...")
    if not songs:
        raise SystemExit(1)
    mine = sharding.shard_songs(len(songs), rank, world)              # song i -> rank i % world, no collective
    n = separate_files(model, songs, args.tar, bool(args.vocal_solo), args.sr, args.songs_per_batch, args.emit_npy, mine)
    print(f"[rank {rank}] wrote {n} files to {args.tar}")


if __name__ == "__main__":
    main()
