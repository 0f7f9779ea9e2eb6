"""GPU: training-step throughput (BASELINE configs[4]): batch 64 per GPU, fused train_step, optional DDP."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import model as svs_model, training  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    net = svs_model.UNet().train().cuda()
    g = torch.Generator(device="cuda").manual_seed(local)
    mix = torch.rand(batch, 1, 512, 128, device="cuda", generator=g)
    voc = mix * torch.rand(batch, 1, 512, 128, device="cuda", generator=g)
    for _ in range(2):
        loss = training.train_step(net, mix, voc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = training.train_step(net, mix, voc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"train_step batch={batch}/GPU x{world}: {ms:.1f} ms/step -> {batch * world / ms * 1e3:.0f} patches/s, "
              f"{3.961 * batch / ms:.2f} TFLOP/s/GPU, loss={float(loss[0]):.5f}")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
