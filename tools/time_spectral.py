"""GPU: STFT / iSTFT+OLA throughput on a corpus-sized ragged batch (BASELINE configs[3] geometry)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import spectral  # noqa: E402


def main():
    n_songs = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 180.0
    n = int(seconds * 8192)
    audio = torch.randn(n_songs * n, device="cuda") * 0.1
    batch = spectral.SongBatch(audio, [n] * n_songs)
    frames = batch.total_frames
    for _ in range(100):                       # ~0.2 s: let the SM clock ramp from idle before timing
        mag, phase, smax = batch.stft()
        wave, peak = batch.istft(mag, phase)
    torch.cuda.synchronize()
    reps = 50
    for name, fn, bytes_per_frame in (("stft", lambda: batch.stft(), 9228), ("istft+ola", lambda: batch.istft(mag, phase), 9228)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = frames * bytes_per_frame / ms / 1e6
        print(f"{name:10s} {ms:8.3f} ms  {frames / ms / 1e3:8.1f} Mframes/s  {gbs:8.1f} GB/s algorithmic "
              f"({gbs / 6551 * 100:.1f}% of measured HBM 6551 GB/s)  audio-s/s {n_songs * seconds / ms * 1e3:.3e}")
    err = (batch.song_wave(wave, 0) - audio[: batch.wave_lengths[0]]).abs().max().item()
    print("round-trip max err", err)


if __name__ == "__main__":
    main()
