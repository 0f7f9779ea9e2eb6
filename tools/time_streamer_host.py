"""GPU: host-side enqueue cost of pipeline.PatchStreamer.run per step (perf_counter around the loop, before the
final synchronisation) next to the device-timed step: is the e2e path bound by the copies or by the Python loop?"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import model as svs_model, pipeline  # noqa: E402


def main():
    batch = 64
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    torch.manual_seed(0)
    net = svs_model.UNet(precision="bf16").eval().cuda()
    hin = [torch.rand(batch, 1, 512, 128).pin_memory() for _ in range(8)]
    hout = [torch.empty(batch, 1, 512, 128).pin_memory() for _ in range(8)]
    seq_in = [hin[i % 8] for i in range(n)]
    seq_out = [hout[i % 8] for i in range(n)]
    st = pipeline.PatchStreamer(net, batch, vocal_solo=True)
    st.run(seq_in[:8], seq_out[:8])
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        st.run(seq_in, seq_out)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"n={n}: host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, wall to completion {1e3 * (t2 - t0) / n:.3f} ms/step "
              f"= {batch * n / (t2 - t0):.0f} patches/s")


if __name__ == "__main__":
    main()
