"""GPU: where a training step's time goes -- graph replay alone, Adam alone, host time of the whole call."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import model as svs_model, training  # noqa: E402


def timed(fn, n=50):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n


def main():
    torch.manual_seed(0)
    net = svs_model.UNet().train().cuda()
    mix = torch.rand(64, 1, 512, 128, device="cuda")
    voc = mix * torch.rand_like(mix)
    for _ in range(3):
        training.train_step(net, mix, voc)
    dev, host = timed(lambda: training.train_step(net, mix, voc))
    print(f"train_step                : {dev:.3f} ms device, {host:.3f} ms host enqueue")
    dev, host = timed(lambda: training.train_step(net, mix, voc, step=False))
    print(f"train_step(step=False)    : {dev:.3f} ms device, {host:.3f} ms host enqueue")
    g = training._get_graph(net, mix, voc, True, 1.0, False)
    dev, host = timed(lambda: g.graphs[0].replay())
    print(f"graph replay alone        : {dev:.3f} ms device, {host:.3f} ms host enqueue")
    dev, host = timed(lambda: net.optim.step())
    print(f"optim.step() alone        : {dev:.3f} ms device, {host:.3f} ms host enqueue")


if __name__ == "__main__":
    main()
