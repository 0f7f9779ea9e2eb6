"""GPU: per-CTA clock64 trace of one tcgen05 conv layer (profiling only)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import _lib, model as svs_model  # noqa: E402


def main():
    layers = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [3]
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    torch.manual_seed(0)
    net = svs_model.UNet(precision="bf16").eval().cuda()
    x = torch.rand(batch, 1, 512, 128, device="cuda")
    out = torch.empty_like(x)
    plan = net.plan()
    iv = _lib.PatchView(x.data_ptr(), None, 512 * 128, 128, 1)
    ov = _lib.PatchView(out.data_ptr(), None, 512 * 128, 128, 1)
    for _ in range(3):
        plan.forward_views(iv, ov, None, batch, 0)
    torch.cuda.synchronize()
    buf = torch.zeros(8 * 1024, dtype=torch.int64, device="cuda")
    lib = _lib.load()
    for li in layers:
        buf.zero_()
        lib.svs_debug_set_trace(buf.data_ptr(), li)
        plan.forward_views(iv, ov, None, batch, 0)        # whole forward: the traced layer runs in sequence
        torch.cuda.synchronize()
        lib.svs_debug_set_trace(None, -1)
        t = buf.view(-1, 8).cpu()
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        names = ["start", "setup", "first_full", "mma_issued", "acc_ready", "epi_done", "exit", "parked_sync"]
        print(f"layer {li}: {t.shape[0]} CTAs; ns since the first CTA started (globaltimer):")
        for i, n in enumerate(names):
            col = t[:, i]
            col = (col[col > 0] - t0).float()
            if col.numel():
                print(f"   {n:11s} median {col.median().item():9.0f}  min {col.min().item():9.0f}  max {col.max().item():9.0f}")


if __name__ == "__main__":
    main()
