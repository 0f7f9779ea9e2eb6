"""GPU (torchrun): NCCL all-reduce time of the flat 39.3 MB gradient buffer."""
import os

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
flat = torch.zeros(9823313, device="cuda")
for _ in range(5):
    dist.all_reduce(flat)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dist.all_reduce(flat)
e1.record()
torch.cuda.synchronize()
if dist.get_rank() == 0:
    ms = e0.elapsed_time(e1) / 20
    print(f"all_reduce 39.3 MB x{dist.get_world_size()}: {ms:.3f} ms  busbw {2 * (dist.get_world_size() - 1) / dist.get_world_size() * 39.3e-3 / ms:.1f} GB/s")
dist.destroy_process_group()
