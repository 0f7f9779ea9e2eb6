"""GPU: throughput of TWO independent 64-patch forwards in flight (two streams, two workspaces, one plan) against one:
do the layers that leave SMs idle (128 CTAs on 148 SMs, one-tile-per-CTA tails) fill up from a second batch?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import _lib, model as svs_model  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    torch.manual_seed(0)
    net = svs_model.UNet(precision="bf16").eval().cuda()
    plan = net.plan()
    pool = 10
    xs = [torch.rand(batch, 1, 512, 128, device="cuda") for _ in range(pool)]
    ys = [torch.empty_like(xs[0]) for _ in range(pool)]
    flags = _lib.FLAG_APPLY_MASK
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]

    def graph_on(stream, idx):
        with torch.cuda.stream(stream):
            for i in idx[:2]:
                plan.forward_dense(xs[i], flags, ys[i])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            with torch.cuda.graph(g, stream=stream):
                for i in idx:
                    plan.forward_dense(xs[i], flags, ys[i])
        torch.cuda.synchronize()
        return g

    g_all = graph_on(streams[0], list(range(pool)))
    g_a = graph_on(streams[0], list(range(0, pool, 2)))
    g_b = graph_on(streams[1], list(range(1, pool, 2)))

    def timed(fn, reps=100):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def one():
        with torch.cuda.stream(streams[0]):
            g_all.replay()

    def two():
        with torch.cuda.stream(streams[0]):
            g_a.replay()
        with torch.cuda.stream(streams[1]):
            g_b.replay()

    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    t1 = timed(one) / pool
    t2 = timed(two) / pool
    print(f"one forward in flight : {t1 * 1e3:.1f} us per {batch} patches = {batch / t1 * 1e3:.0f} patches/s")
    print(f"two forwards in flight: {t2 * 1e3:.1f} us per {batch} patches = {batch / t2 * 1e3:.0f} patches/s")


if __name__ == "__main__":
    main()
