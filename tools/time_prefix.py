"""GPU: per-layer device time INSIDE the forward, the way bench.py reports `layer_us`: CUDA graphs of the prefixes
conv1..layer over a pool of inputs larger than L2; a layer's figure is the difference of consecutive prefixes.

    python tools/time_prefix.py [batch] [precision] [--total-only]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import _lib, model as svs_model  # noqa: E402

NAMES = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "deconv1", "deconv2", "deconv3", "deconv4",
         "deconv5", "deconv6"]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    batch = int(args[0]) if len(args) > 0 else 64
    prec = args[1] if len(args) > 1 else "bf16"
    total_only = "--total-only" in sys.argv
    torch.manual_seed(0)
    net = svs_model.UNet(precision=prec).eval().cuda()
    plan = net.plan()
    pool = max(2, (160 << 20) // (batch * 512 * 128 * 4) + 1)
    xs = [torch.rand(batch, 1, 512, 128, device="cuda") for _ in range(pool)]
    ys = [torch.empty_like(xs[0]) for _ in range(pool)]
    flags = _lib.FLAG_APPLY_MASK
    side = torch.cuda.Stream()

    def graph(last):
        def one(i):
            iv = _lib.PatchView(xs[i].data_ptr(), None, 512 * 128, 128, 1)
            ov = _lib.PatchView(ys[i].data_ptr(), None, 512 * 128, 128, 1)
            plan.forward_views(iv, ov, None, batch, flags, 0, last)
        for i in range(3):
            one(i % pool)
        torch.cuda.synchronize()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for i in range(pool):
                    one(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        return g

    def time_graph(g, reps):
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * pool)

    reps = max(5, int(0.2 / (pool * 0.25e-3 * batch / 64)))
    if total_only:
        ms = min(time_graph(graph(11), reps) for _ in range(3))
        print(f"batch={batch} prec={prec} {ms * 1e3:.1f} us/forward  {batch / ms * 1e3:.0f} patches/s")
        return
    pre = [time_graph(graph(li), reps) for li in range(12)]
    prev = 0.0
    for li, name in enumerate(NAMES):
        print(f"{name:8s} {(pre[li] - prev) * 1e3:7.1f} us")
        prev = pre[li]
    print(f"batch={batch} prec={prec} total {pre[11] * 1e3:.1f} us/forward  {batch / pre[11] * 1e3:.0f} patches/s")


if __name__ == "__main__":
    main()
