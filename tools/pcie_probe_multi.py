"""Developer tool (N GPUs, torchrun): aggregate device->host and host->device bandwidth of the box when every rank copies at
once (24.9 MB images down, 5.6 MB cloud slices up) -- what the e2e leg of the bench can get at N GPUs."""
import json
import os

import torch
import torch.distributed as dist

dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
img_d = torch.empty(24883200, dtype=torch.uint8, device=dev)
img_h = [torch.empty(24883200, dtype=torch.uint8).pin_memory() for _ in range(3)]
up_h = torch.empty(44798092 // world, dtype=torch.uint8).pin_memory()
up_d = torch.empty(44798092 // world, dtype=torch.uint8, device=dev)
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def run(down, up, n=60):
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        if up:
            with torch.cuda.stream(s_up):
                up_d.copy_(up_h, non_blocking=True)
        if down:
            with torch.cuda.stream(s_dn):
                img_h[i % 3].copy_(img_d, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_up)
    torch.cuda.current_stream().wait_stream(s_dn)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


for down, up in ((True, False), (False, True), (True, True)):
    run(down, up, 5)
    ms = run(down, up)
    if rank == 0:
        print(json.dumps({"ranks": world, "down": down, "up": up, "ms_per_step": round(ms, 4), "steps_per_s_per_rank": round(1e3 / ms, 1),
                          "aggregate_d2h_GBps": round(world * 24.8832 / ms, 1) if down else 0,
                          "aggregate_h2d_GBps": round(44.798 / ms, 1) if up else 0}))
dist.destroy_process_group()
