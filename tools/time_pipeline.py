"""GPU: full-song pipeline throughput (BASELINE configs[2]/[3]): audio -> STFT -> UNet mask -> iSTFT, device resident."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import model as svs_model, pipeline, spectral  # noqa: E402


def main():
    n_songs = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 180.0
    prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    max_batch = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    n = int(seconds * 8192)
    torch.manual_seed(0)
    net = svs_model.UNet(precision=prec).eval().cuda()
    audio = torch.randn(n_songs * n, device="cuda") * 0.1
    batch = spectral.SongBatch(audio, [n] * n_songs)
    sep = pipeline.Separator(net, max_batch=max_batch)
    for _ in range(2):
        wave, peak = sep.separate_batch(batch)
    torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        wave, peak = sep.separate_batch(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    wall = (time.perf_counter() - t0) / reps * 1e3
    patches = n_songs * (batch.frames[0] // 128 + 1)
    print(f"{n_songs} songs x {seconds:.0f}s ({patches} patches, {batch.total_frames} frames) [{prec}, max_batch {max_batch}]: {ms:.2f} ms device "
          f"({wall:.2f} ms wall) -> {n_songs * seconds / ms * 1e3:.3e} audio-s/s, {patches / ms * 1e3:.0f} patches/s")


if __name__ == "__main__":
    main()
