"""GPU: device time of a contiguous layer range of the UNet forward, back to back (warm L2, PDL chain)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import _lib, model as svs_model  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    ranges = [tuple(int(v) for v in a.split("-")) for a in sys.argv[3:]] or [(0, 11), (4, 7), (4, 4), (5, 5), (6, 6), (7, 7)]
    torch.manual_seed(0)
    net = svs_model.UNet(precision=prec).eval().cuda()
    x = torch.rand(batch, 1, 512, 128, device="cuda")
    out = torch.empty_like(x)
    plan = net.plan()
    iv = _lib.PatchView(x.data_ptr(), None, 512 * 128, 128, 1)
    ov = _lib.PatchView(out.data_ptr(), None, 512 * 128, 128, 1)
    for _ in range(5):
        plan.forward_views(iv, ov, None, batch, 0)
    torch.cuda.synchronize()
    for first, last in ranges:
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                plan.forward_views(iv, ov, None, batch, 0, first, last)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 50 * 1e3)
        print(f"layers {first:2d}..{last:2d}: {best:8.1f} us  (batch {batch}, {prec})", flush=True)


if __name__ == "__main__":
    main()
