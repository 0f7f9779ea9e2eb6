"""GPU: where the device time of Separator.separate_batch goes (events between the stages)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import _lib, model as svs_model, pipeline, spectral  # noqa: E402


def main():
    n_songs, seconds, max_batch = 150, 180.0, int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n = int(seconds * 8192)
    torch.manual_seed(0)
    net = svs_model.UNet(precision="bf16").eval().cuda()
    audio = torch.randn(n_songs * n, device="cuda") * 0.1
    batch = spectral.SongBatch(audio, [n] * n_songs)
    plan = net.plan()
    names = ["stft", "table+norm idx", "gather", "unet", "scatter", "istft+peak"]
    acc = [0.0] * len(names)
    reps = 6
    for rep in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
        spans = []
        ev[0].record()
        mag, phase, smax = batch.stft()
        ev[1].record()
        offs, valid, song = pipeline.patch_table(batch.frames, batch.frame_off_host)
        d_off = torch.from_numpy(offs).cuda(); d_valid = torch.from_numpy(valid).cuda()
        d_norm = smax[torch.from_numpy(song).cuda().long()]
        out_mag = torch.empty_like(mag)
        ev[2].record()
        g = u = s = 0.0
        marks = []
        for a in range(0, len(offs), max_batch):
            b = min(len(offs), a + max_batch)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
            x = _lib.patches_gather_raw(mag, d_off[a:b], d_valid[a:b], d_norm[a:b])
            e[1].record()
            y = plan.forward_dense(x, _lib.FLAG_APPLY_MASK)
            e[2].record()
            _lib.patches_scatter_raw(y, d_off[a:b], d_valid[a:b], out_mag, dc_zero=True)
            e[3].record()
            marks.append(e)
        ev[3].record()
        wave, peak = batch.istft(out_mag, phase, peak_normalize=True)
        ev[4].record()
        torch.cuda.synchronize()
        if rep == 0:
            continue
        acc[0] += ev[0].elapsed_time(ev[1]); acc[1] += ev[1].elapsed_time(ev[2])
        for e in marks:
            acc[2] += e[0].elapsed_time(e[1]); acc[3] += e[1].elapsed_time(e[2]); acc[4] += e[2].elapsed_time(e[3])
        acc[5] += ev[3].elapsed_time(ev[4])
    for nme, v in zip(names, acc):
        print(f"{nme:16s} {v / (reps - 1):8.3f} ms")
    print(f"{'sum':16s} {sum(acc) / (reps - 1):8.3f} ms")


if __name__ == "__main__":
    main()
