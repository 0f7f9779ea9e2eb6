"""GPU: per-CTA clock64 trace of one zero-copy tcgen05 layer, run inside the full forward (profiling only)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import _lib, model as svs_model  # noqa: E402


def main():
    layers = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [10]
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
    buf = torch.zeros(64 * 1024, dtype=torch.int64, device="cuda")
    lib = _lib.load()
    for li in layers:
        buf.zero_()
        lib.svs_debug_set_trace(buf.data_ptr(), li)
        plan.forward_views(iv, ov, None, batch, 0)
        torch.cuda.synchronize()
        lib.svs_debug_set_trace(None, -1)
        t = buf.view(-1, 64).cpu()
        t = t[t[:, 0] > 0]
        med = lambda col: int((t[:, col] - t[:, 0])[t[:, col] > 0].float().median().item()) if (t[:, col] > 0).any() else -1
        print(f"layer {li}: {t.shape[0]} CTAs; median cycles since CTA start")
        print("   exit           ", med(1))
        print("   A issued  job0-7", [med(8 + j) for j in range(8)])
        print("   A landed  job0-7", [med(16 + j) for j in range(8)])
        print("   MMA done issuing tile0-7", [med(24 + j) for j in range(8)])
        print("   epilogue start tile0-7 ", [med(32 + j) for j in range(8)])
        print("   epilogue end   tile0-7 ", [med(40 + j) for j in range(8)])
        if (t[:, 48] > 0).any():                       # TMA-store epilogue: stamps inside the first two 32-column steps
            print("   store step 0 [ld issue, ld done, smem written, store ring free, barrier, TMA issued]", [med(48 + j) for j in range(6)])
            print("   store step 1", [med(54 + j) for j in range(6)])


if __name__ == "__main__":
    main()
