#!/usr/bin/env python3
"""Aggregate pinned host->device bandwidth with every GPU of the box copying at once (one process per GPU under torchrun):
the ceiling of bench.py's end-to-end leg at N GPUs."""
import os, json, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 604372992
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
t = torch.tensor([dt], device="cuda", dtype=torch.float64)
if world > 1:
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    ts = [float(x.item()) for x in allt]
else:
    ts = [dt]
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_copy": n, "ms_per_copy_per_rank": [x * 1e3 for x in ts], "aggregate_GBps": sum(n / x for x in ts) / 1e9,
                      "e2e_ceiling_Gbit_s": world * 16384 * 6144 / max(ts) / 1e9}))
if world > 1:
    dist.destroy_process_group()
