"""GPU (optionally torchrun, N ranks): the copy-only ceiling of the host-buffer paths — pinned host -> device and
device -> pinned host at the same time, no kernels — for the two transfer sizes the pipelines use.  Run it with the
same N as bench.py to see what the box's PCIe / host-memory path can carry when every rank copies at once:

  python tools/time_copies.py
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/time_copies.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200 import pipeline  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = pipeline.bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank() if world > 1 else 0
    for label, nbytes, steps in (("64-patch batch (16.8 MB each way)", 64 * 512 * 128 * 4, 200),
                                 ("10-song chunk, fp32 (59 MB each way)", 10 * 180 * 8192 * 4, 60),
                                 ("10-song chunk, PCM_16 (29 MB each way)", 10 * 180 * 8192 * 2, 100)):
        hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        c = pipeline.CopyCeiling(dev, nbytes)
        c.run(hin, hout, 5)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        c.run(hin, hout, steps)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
        if rank == 0:
            print(f"N={world} {label}: {ms:.3f} ms per step (max over ranks) -> {nbytes / ms / 1e6:.1f} GB/s per rank per "
                  f"direction, {world * 2 * nbytes / ms / 1e6:.0f} GB/s aggregate both directions; numa={numa.get('numa_node')}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
