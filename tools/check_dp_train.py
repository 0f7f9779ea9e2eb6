"""GPU (torchrun, N ranks): the data-parallel training step — staged backward with the decoder's gradient all-reduce
running under the encoder's backward — gives the mean of the ranks' local gradients, and its timing."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import model as svs_model, training  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    torch.manual_seed(0)
    net = svs_model.UNet().train().cuda()
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0                                                  # deterministic comparison
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    mix = torch.rand(batch, 1, 512, 128, device="cuda", generator=g)
    voc = mix * torch.rand(batch, 1, 512, 128, device="cuda", generator=g)
    bufs = [b.clone() for b in net.buffers()]
    training.train_step(net, mix, voc, step=False, sync_grads=False)
    local_grad = net._flat_grad.clone()
    for b, v in zip(net.buffers(), bufs):
        b.copy_(v)
    gathered = [torch.empty_like(local_grad) for _ in range(world)]
    dist.all_gather(gathered, local_grad)
    want = torch.stack(gathered).double().mean(0)
    for use_graph in (False, True):
        for b, v in zip(net.buffers(), bufs):
            b.copy_(v)
        training.train_step(net, mix, voc, step=False, sync_grads=True, use_graph=use_graph)
        got = net._flat_grad.double()
        err = float((got - want).norm() / want.norm())
        if rank == 0:
            print(f"world {world} graph={use_graph}: |synced - mean(local)| / |mean| = {err:.2e}")
        assert err < 1e-6, err
    for sync in (True, False):
        for _ in range(3):
            training.train_step(net, mix, voc, sync_grads=sync)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            training.train_step(net, mix, voc, sync_grads=sync)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world {world} batch {batch}/GPU sync_grads={sync}: {float(t):.3f} ms/step -> {batch * world / float(t) * 1e3:.0f} patches/s")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
