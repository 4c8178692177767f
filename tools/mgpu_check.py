"""torchrun --nproc-per-node N tools/mgpu_check.py : NCCL-backed checks of distributed.py on real GPUs."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import poppunk_mod_b200 as pk
from oracle import workloads as wl
from poppunk_mod_b200 import distributed as D

obs = wl.synthetic_target(50_000)
sims = [wl.sim_batch_member(s, n=40_000) for s in range(13)]
got = D.batch_discrepancy_sharded(obs, sims, 0.25, 0.0, False, resolution=128)
want = pk.KDE_JS_divergence_batch(obs, sims, 0.25, 0.0, False, resolution=128)
assert np.array_equal(got, want), (got, want)

# config C5 shape at reduced and full scale: one evaluation, node rows split over the ranks
for n, R in ((1_000_000, 512), (int(os.environ.get("C5_N", 10_000_000)), 1024)):
    pool = pk._native.PinnedPool()                      # page-locked inputs: H2D at PCIe speed
    a = pool.empty((n, 2)); a[:] = wl.synthetic_target(n, seed=11)
    b = pool.empty((n, 2)); b[:] = wl.sim_S(5, 1800, 0.36, n=n)
    D.row_sharded_discrepancy(a, b, 0.25, 0.0, resolution=R)          # warm-up (scratch allocation)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    js = D.row_sharded_discrepancy(a, b, 0.25, 0.0, resolution=R)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    # the same evaluation with the TABLE sharded too: each rank uploads 1/world of the rows, NCCL all-gather
    ca = [len(a) * r // world for r in range(world + 1)]; cb = [len(b) * r // world for r in range(world + 1)]
    sa, sb = a[ca[rank]:ca[rank + 1]], b[cb[rank]:cb[rank + 1]]
    D.row_sharded_discrepancy(sa, sb, 0.25, 0.0, resolution=R, sliced=True)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    js2 = D.row_sharded_discrepancy(sa, sb, 0.25, 0.0, resolution=R, sliced=True)
    torch.cuda.synchronize(); dist.barrier(); dt2 = time.perf_counter() - t0
    assert js2 == js, (js2, js)
    if rank == 0:
        print(f"   sliced inputs + NCCL all-gather: {dt2*1e3:.1f} ms", flush=True)
        pk.KDE_JS_divergence(a, b, 0.25, 0.0, resolution=R)
        t0 = time.perf_counter(); full = pk.KDE_JS_divergence(a, b, 0.25, 0.0, resolution=R); d1 = time.perf_counter() - t0
        print(f"C5-like n={n} R={R}: sharded over {world} = {js!r} in {dt*1e3:.1f} ms ; single GPU = {full!r} in {d1*1e3:.1f} ms ; rel {abs(js-full)/full:.2e}", flush=True)
        assert abs(js - full) <= 1e-12 * full
if rank == 0:
    print("mgpu_check ok", flush=True)
dist.destroy_process_group()
