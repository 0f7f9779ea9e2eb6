"""GPU: per-layer device time of the UNet forward (CUDA events around svs_unet_forward_layers)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import _lib, model as svs_model  # noqa: E402

NAMES = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "deconv1", "deconv2", "deconv3", "deconv4",
         "deconv5", "deconv6"]
# nominal GFLOP per patch per layer (2 * MACs incl. padded taps), SURVEY.md section 8(a)
GF = [0.0131, 0.1049, 0.1049, 0.1049, 0.1049, 0.1049, 0.1049, 0.2097, 0.2097, 0.2097, 0.2097, 0.0262]


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    torch.manual_seed(0)
    net = svs_model.UNet(precision=prec).eval().cuda()
    x = torch.rand(batch, 1, 512, 128, device="cuda")
    out = torch.empty_like(x)
    plan = net.plan()
    iv = _lib.PatchView(x.data_ptr(), None, 512 * 128, 128, 1)
    ov = _lib.PatchView(out.data_ptr(), None, 512 * 128, 128, 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        plan.forward_views(iv, ov, None, batch, 0)
    torch.cuda.synchronize()
    total = 0.0
    for li, name in enumerate(NAMES):
        ts = []
        for _ in range(iters):
            flush.zero_()                                   # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.forward_views(iv, ov, None, batch, 0, li, li)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        total += ms
        print(f"{name:8s} {ms * 1e3:8.1f} us   {GF[li] * batch / ms:8.1f} TFLOP/s(nominal)")
    print(f"sum {total * 1e3:.1f} us (cold L2, each layer alone)")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        plan.forward_views(iv, ov, None, batch, 0)
    e1.record()
    torch.cuda.synchronize()
    print(f"full forward back-to-back: {e0.elapsed_time(e1) / iters * 1e3:.1f} us")


if __name__ == "__main__":
    main()
