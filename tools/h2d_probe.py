"""Pinned host -> device copy rate at 1/2/4/8 concurrent ranks: the ceiling of bench.py's end-to-end number on a host.

    python tools/h2d_probe.py                                                   # one rank
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py

Every rank copies from its own page-locked buffer to its own GPU at the same time (barrier before each size); per
size it prints the per-rank rate (min / mean over ranks) and the aggregate.  1.6 MB is one 100k-pair simulated set,
25.6 MB a wave of 16 uploaded with one copy (PinnedPool.empty_batch)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poppunk_mod_b200 as pk

numa = pk._native.bind_host_to_device(local) if world > 1 else -1
x = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64).pin_memory()
d = torch.empty_like(x, device="cuda")
rows = []
for mb in (1.6, 6.4, 25.6, 102.4, 256.0):
    k = int(mb * 1e6 // 8)
    reps = max(3, int(1024 // mb))
    for _ in range(3):
        d[:k].copy_(x[:k], non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        d[:k].copy_(x[:k], non_blocking=True)
    torch.cuda.synchronize()
    gbs = torch.tensor([k * 8 * reps / (time.perf_counter() - t0) / 1e9], dtype=torch.float64, device="cuda")
    if world > 1:
        allv = [torch.empty_like(gbs) for _ in range(world)]
        dist.all_gather(allv, gbs)
        vals = [float(v.item()) for v in allv]
    else:
        vals = [float(gbs.item())]
    rows.append({"copy_mb": mb, "per_rank_gbs_min": min(vals), "per_rank_gbs_mean": sum(vals) / len(vals),
                 "aggregate_gbs": sum(vals)})
if rank == 0:
    print(json.dumps({"ranks": world, "numa_node_rank0": numa, "host_cpus": os.cpu_count(), "h2d_pinned": rows}))
if world > 1:
    dist.destroy_process_group()
