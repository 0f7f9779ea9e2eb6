"""GPU: time UNet forward at a given batch / precision with CUDA events (device-resident input)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import model as svs_model  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    torch.manual_seed(0)
    net = svs_model.UNet(precision=prec).eval().cuda()
    x = torch.rand(batch, 1, 512, 128, device="cuda")
    out = torch.empty_like(x)
    plan = net.plan()
    for _ in range(3):
        plan.forward_dense(x, 0, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        plan.forward_dense(x, 0, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"batch={batch} prec={prec} {ms:.4f} ms/forward  {batch / ms * 1e3:.0f} patches/s  "
          f"{batch * 1.3247e9 / ms / 1e9:.1f} TFLOP/s(exact)  launches={plan.launch_count(batch)}")


if __name__ == "__main__":
    main()
