"""GPU experiment: a 64-patch batch as two 32-patch half-batches on two streams (layer pipelining across halves)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import model as svs_model  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    ways = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    torch.manual_seed(0)
    nets = [svs_model.UNet(precision="bf16").eval().cuda() for _ in range(ways)]
    for n in nets[1:]:
        n.load_state_dict(nets[0].state_dict())
    plans = [n.plan() for n in nets]
    x = torch.rand(batch, 1, 512, 128, device="cuda")
    out = torch.empty_like(x)
    h = batch // ways
    streams = [torch.cuda.Stream() for _ in range(ways)]
    main_s = torch.cuda.current_stream()

    def step():
        ev = torch.cuda.Event()
        ev.record(main_s)
        for i, (p, s) in enumerate(zip(plans, streams)):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                p.forward_dense(x[i * h:(i + 1) * h], 0, out[i * h:(i + 1) * h])
            e2 = torch.cuda.Event()
            e2.record(s)
            main_s.wait_event(e2)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(main_s)
    with torch.cuda.stream(side):
        main_save = main_s
        with torch.cuda.graph(g, stream=side):
            ev = torch.cuda.Event()
            ev.record(side)
            for i, (p, s) in enumerate(zip(plans, streams)):
                s.wait_event(ev)
                with torch.cuda.stream(s):
                    p.forward_dense(x[i * h:(i + 1) * h], 0, out[i * h:(i + 1) * h])
                e2 = torch.cuda.Event()
                e2.record(s)
                side.wait_event(e2)
    main_s.wait_stream(side)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 100 * 1e3)
    print(f"batch {batch} as {ways} x {h} on {ways} streams (graph replay): {best:.1f} us -> {batch / best * 1e6:.0f} patches/s")


if __name__ == "__main__":
    main()
