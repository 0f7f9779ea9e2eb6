"""GPU: host-to-host corpus throughput of pipeline.SongStreamer for a few chunk sizes / stream counts."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import model as svs_model, pipeline  # noqa: E402


def main():
    n_songs, n = 150, int(180.0 * 8192)
    torch.manual_seed(0)
    net = svs_model.UNet(precision="bf16").eval().cuda()
    host_audio = (torch.randn(n_songs * n) * 0.1).pin_memory()
    host_wave = torch.empty(n_songs * 768 * (n // 768), dtype=torch.float32).pin_memory()
    lengths = [n] * n_songs
    for chunk, streams in ((15, 2), (15, 4), (10, 4), (6, 4), (10, 3), (25, 4)):
        st = pipeline.SongStreamer(net, songs_per_chunk=chunk, n_streams=streams)
        for _ in range(2):
            st.run(host_audio, lengths, host_wave)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            st.run(host_audio, lengths, host_wave)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        print(f"chunk {chunk:3d} songs, {streams} streams: {ms:7.2f} ms/corpus -> {n_songs * 180.0 / ms * 1e3:.3e} audio-s/s "
              f"({2 * host_audio.numel() * 4 / ms / 1e6:.1f} GB/s PCIe both ways)")


if __name__ == "__main__":
    main()
