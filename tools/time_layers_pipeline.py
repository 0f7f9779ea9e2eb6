"""GPU: per-layer device time of conv1 / deconv6 through the frame-major patch views of the song pipeline."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import _lib, model as svs_model, pipeline  # noqa: E402


def main():
    torch.manual_seed(0)
    net = svs_model.UNet(precision="bf16").eval().cuda()
    plan = net.plan()
    t_song = 1921
    songs = 4
    mag = torch.rand(songs * t_song, 513, device="cuda")
    out = torch.zeros_like(mag)
    offs, valid, _ = pipeline.patch_table([t_song] * songs, np.arange(songs + 1) * t_song)
    d_off = torch.from_numpy(offs).cuda()
    d_valid = torch.from_numpy(valid).cuda()
    n = 64
    iv = _lib.PatchView(mag.data_ptr(), d_off.data_ptr(), 0, 1, 513)
    ov = _lib.PatchView(out.data_ptr(), d_off.data_ptr(), 0, 1, 513)
    for _ in range(3):
        plan.forward_views(iv, ov, d_valid, n, 1)
    torch.cuda.synchronize()
    for li in (0, 11):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            plan.forward_views(iv, ov, d_valid, n, 1, li, li)
        e1.record()
        torch.cuda.synchronize()
        print(f"pipeline-view layer {li}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per 64 patches")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        plan.forward_views(iv, ov, d_valid, n, 1)
    e1.record()
    torch.cuda.synchronize()
    print(f"pipeline-view full forward: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per 64 patches")


if __name__ == "__main__":
    main()
