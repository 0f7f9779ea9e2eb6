"""GPU: one small invocation of every hand-written kernel family, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck); results are compared with nothing here — the parity tests do that.

  compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import _lib, model as svs_model, pipeline, resample, spectral, synth, training  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.manual_seed(0)
    if which in ("all", "spectral"):
        songs = [synth.synth_song(4.0, seed=1)[0], synth.synth_song(2.3, seed=2)[0]]
        batch = spectral.SongBatch.from_audio(songs)
        mag, phase, smax = batch.stft()                                 # K1 stft_mag_phase_kernel
        batch.istft(mag, phase, peak_normalize=True)                    # K2 istft_ola_kernel + normalise
        batch.istft(mag, phase, peak_normalize=True, pcm16=True)
        resample.resample(np.random.default_rng(0).standard_normal((4410, 2)).astype(np.float32), 44100, 8192)
        torch.cuda.synchronize()
        print("spectral ok")
    if which in ("all", "unet"):
        for prec in ("bf16", "tf32"):
            net = svs_model.UNet(precision=prec).eval().cuda()
            x = torch.rand(8, 1, 512, 128, device="cuda")               # 8 patches: cluster split-K on the deep layers
            with torch.no_grad():
                net.separate(x)                                         # conv1_zc, zc_conv, tc_conv_ck / tc_conv, deconv6_tc
            sep = pipeline.Separator(net, max_batch=4)
            sep.separate([synth.synth_song(13.0, seed=3)[0]])           # patch gather / scatter
            torch.cuda.synchronize()
            print("unet", prec, "ok")
    if which in ("all", "train"):
        net = svs_model.UNet().train().cuda()
        mix = torch.rand(2, 1, 512, 128, device="cuda")
        voc = mix * torch.rand_like(mix)
        for prec in ("tf32", "fp32"):
            net.train_precision = prec
            training.train_step(net, mix, voc, use_graph=False)        # tc fwd / dgrad, wgrad_tc, BN kernels, edge wgrad
        torch.cuda.synchronize()
        print("train ok")


if __name__ == "__main__":
    main()
