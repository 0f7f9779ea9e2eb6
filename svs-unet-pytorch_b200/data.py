"""Drop-in for reference data.py (same flags, same directory / .npy conventions) on the GPU kernels.

    python scripts/data.py --src <songs> --tar <specs> --direction to_spec
    python scripts/data.py --src <specs> --phase <specs/mixture> --tar <wavs> --direction to_wave

to_spec (reference data.py:46-112): per song folder, mixture.wav / vocals.wav -> STFT -> magphase ->
divide by the MIXTURE's max magnitude -> ``<tar>/{mixture,vocal}/NNNN_<song>_{spec,phase}.npy``
(float32 / complex64, (513, T), Fortran order exactly as librosa + np.save produce).
to_wave (reference data.py:117-169): spec * phase -> iSTFT -> 0.9 peak -> PCM_16 wav."""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from . import audio_io, spectral
from .config import HOP_SIZE, SAMPLE_RATE, WINDOW_SIZE, num2str

TRACK_MAP = {"mixture.wav": "mixture", "vocals.wav": "vocal"}


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--src", type=str, required=True)
    p.add_argument("--tar", type=str, required=True)
    p.add_argument("--phase", type=str, default="-1")
    p.add_argument("--win_size", type=int, default=WINDOW_SIZE)
    p.add_argument("--hop_size", type=int, default=HOP_SIZE)
    p.add_argument("--sr", type=int, default=SAMPLE_RATE)
    p.add_argument("--direction", default="to_spec", choices=["to_spec", "to_wave"])
    return p


def song_to_spec(y_mix: np.ndarray, tracks: dict):
    """One song: {"mixture": y, "vocal": y, ...} -> {name: (spec f32 (513,T) F-order, phase c64 (513,T))}.
    Every stem is length-aligned to the mixture (data.py:97-98) and divided by the mixture's max."""
    names = list(tracks)
    aligned = []
    for n in names:
        y = tracks[n]
        y = y[: len(y_mix)] if len(y) > len(y_mix) else np.pad(y, (0, len(y_mix) - len(y)))
        aligned.append(np.ascontiguousarray(y, dtype=np.float32))
    batch = spectral.SongBatch.from_audio([np.ascontiguousarray(y_mix, dtype=np.float32)] + aligned)
    mag, phase, smax = batch.stft()
    norm = smax[0:1].expand(batch.n_songs).contiguous()              # the MIXTURE's max for every stem
    batch.normalize(mag, norm)
    out = {}
    for i, n in enumerate(names, start=1):
        a, b = int(batch.frame_off_host[i]), int(batch.frame_off_host[i + 1])
        spec = mag[a:b].cpu().numpy().T                              # (513, T), Fortran order
        ph = torch.view_as_complex(phase[a:b]).cpu().numpy().T
        out[n] = (spec, ph)
    return out


def to_spec(args):
    spectral._check_geometry(args.win_size, args.hop_size)
    os.makedirs(args.tar, exist_ok=True)
    for folder in TRACK_MAP.values():
        os.makedirs(os.path.join(args.tar, folder), exist_ok=True)
    songs = sorted(d for d in os.listdir(args.src) if os.path.isdir(os.path.join(args.src, d)))
    print(f"found {len(songs)} song folders in {args.src}")
    if not songs:
        raise SystemExit(1)
    for idx, song in enumerate(songs):
        path = os.path.join(args.src, song)
        mix_path = os.path.join(path, "mixture.wav")
        if not os.path.exists(mix_path):
            continue
        try:
            y_mix = audio_io.load(mix_path, sr=args.sr, mono=True)
            tracks = {}
            for wav, folder in TRACK_MAP.items():
                tp = os.path.join(path, wav)
                if os.path.exists(tp):
                    tracks[folder] = y_mix if wav == "mixture.wav" else audio_io.load(tp, sr=args.sr, mono=True)
            base = f"{num2str(idx)}_{song}"
            for folder, (spec, ph) in song_to_spec(y_mix, tracks).items():
                np.save(os.path.join(args.tar, folder, f"{base}_spec.npy"), spec)
                np.save(os.path.join(args.tar, folder, f"{base}_phase.npy"), ph)
        except Exception as e:                                       # per-song isolation, data.py:111-112
            print(f"Error processing {song}: {e}")


def spec_to_wave(mag: np.ndarray, phase: np.ndarray) -> np.ndarray:
    """reference data.py:151-164 for one song."""
    t = min(mag.shape[1], phase.shape[1])
    dev = spectral._device()
    mag_tf = torch.from_numpy(np.ascontiguousarray(mag[:, :t].T, dtype=np.float32)).to(dev)
    ph = np.ascontiguousarray(phase[:, :t].T.astype(np.complex64))
    ph_tf = torch.view_as_real(torch.from_numpy(ph)).to(dev)
    batch = spectral.SongBatch(torch.zeros(1, dtype=torch.float32, device=dev), [HOP_SIZE * (t - 1)])
    wave, _ = batch.istft(mag_tf, ph_tf, peak_normalize=True)
    return wave.cpu().numpy()


def to_wave(args):
    if args.phase == "-1":
        raise Exception("--phase is required for to_wave")
    spectral._check_geometry(args.win_size, args.hop_size)
    os.makedirs(args.tar, exist_ok=True)
    files = sorted(f for f in os.listdir(args.src) if f.endswith("_spec.npy"))
    print(f"restoring {len(files)} files...")
    for name in files:
        try:
            mag = np.load(os.path.join(args.src, name))
            pname = name.replace("_spec.npy", "_phase.npy")
            phase = None
            for cand in (os.path.join(args.phase, pname), os.path.join(args.phase, "mixture", pname)):
                if os.path.exists(cand):
                    phase = np.load(cand)
                    break
            if phase is None:                                        # data.py:148 random-phase fallback
                phase = np.exp(2j * np.pi * np.random.rand(*mag.shape))
            y = spec_to_wave(mag, phase)
            audio_io.write_wav_pcm16(os.path.join(args.tar, name.replace("_spec.npy", ".wav")), y, args.sr)
        except Exception as e:
            print(f"restore failed {name}: {e}")


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.direction == "to_spec":
        to_spec(args)
    else:
        to_wave(args)


if __name__ == "__main__":
    main()
