"""GPU diagnostic (test infrastructure, lives under tests/ because it uses the oracle): per-layer error of the CUDA
UNet against the CPU oracle.   python tests/diag_unet.py [batch] [fp32,bf16,tf32]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_oracle  # noqa: E402
from svs_unet_pytorch_b200 import model as svs_model  # noqa: E402

NAMES = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "deconv1", "deconv2", "deconv3", "deconv4", "deconv5"]


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    precisions = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp32", "bf16", "tf32"]
    torch.manual_seed(0)
    net = svs_model.UNet().eval()
    g = torch.Generator().manual_seed(1)
    x = torch.rand(batch, 1, 512, 128, generator=g)
    with torch.no_grad():
        ref_mask, ref_acts = unet_oracle.unet_forward(net.state_dict(), x, return_activations=True)
    net = net.cuda()
    for prec in precisions:
        net.precision = prec
        t0 = time.time()
        with torch.no_grad():
            mask = net(x.cuda())
        torch.cuda.synchronize()
        print(f"[{prec}] forward ok in {time.time() - t0:.3f}s; launches={net.plan().launch_count(batch)}", flush=True)
        plan = net.plan()
        for i, n in enumerate(NAMES):
            a = plan.read_activation(i, batch).cpu()
            r = ref_acts[n]
            err = (a - r).abs().max().item()
            print(f"  {n:8s} max|ref|={r.abs().max().item():.4f} max_err={err:.3e} rel={err / max(r.abs().max().item(), 1e-9):.3e}",
                  flush=True)
        print(f"  mask     max_err={(mask.cpu() - ref_mask).abs().max().item():.3e}", flush=True)


if __name__ == "__main__":
    main()
