"""Multi-GPU check over NCCL (one process per GPU):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/mgpu_check.py
MGPU_SMALL=1 keeps it to the small parity cases (tests/test_nccl_gpu.py); otherwise config C5 is timed too."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import poppunk_mod_b200 as pk
from oracle import kde_oracle as ko, workloads as wl
from poppunk_mod_b200 import distributed as D

cut = lambda x: x[len(x) * rank // world: len(x) * (rank + 1) // world]

# batches: items dealt to ranks, scalars gathered
obs = wl.synthetic_target(50_000)
sims = [wl.sim_batch_member(s, n=40_000) for s in range(13)]
got = D.batch_discrepancy_sharded(obs, sims, 0.25, 0.0, False, resolution=128)
want = pk.KDE_JS_divergence_batch(obs, sims, 0.25, 0.0, False, resolution=128)
assert np.array_equal(got, want), (got, want)

# one evaluation, points sharded (band mode) against the single GPU and the oracle; both culled kernels
a, b = wl.synthetic_target(200_000, seed=3), wl.sim_S(8, 2100, 0.31, n=150_000)
for algo in ("binned", "binned32"):
    for R in (100, 256):
        info = {}
        js = D.band_sharded_discrepancy(cut(a), cut(b), 0.25, 0.0, resolution=R, algo=algo, info=info)
        full = pk.KDE_JS_divergence(a, b, 0.25, 0.0, resolution=R, algo=algo)
        assert abs(js - full) <= 1e-12 * full, (algo, R, js, full)
        # the exchange over peer memory and the NCCL all-to-all feed the merge the same points in the same order
        os.environ["KDEJS_BAND_PEER"] = "0"
        info_nccl = {}
        js_nccl = D.band_sharded_discrepancy(cut(a), cut(b), 0.25, 0.0, resolution=R, algo=algo, info=info_nccl)
        del os.environ["KDEJS_BAND_PEER"]
        assert js_nccl == js and info_nccl["exchange"] == "nccl", (js, js_nccl, info_nccl)
        assert world == 1 or os.environ.get("EXPECT_PEER", "0") != "1" or info["exchange"] == "peer memory", info
        if rank == 0:
            o = ko.discrepancy(a, b, 0.25, 0.0, resolution=R)
            assert abs(js - o) <= (1e-10 if algo == "binned" else 2e-7) * o, (algo, R, js, o)
            print(f"band mode {algo} R={R}: {js!r} single {full!r} oracle {o!r} band {info}", flush=True)
# the older replicate-and-sort variant still agrees
js_old = D.row_sharded_discrepancy(cut(a), cut(b), 0.25, 0.0, resolution=256, sliced=True, gather_all=True)
assert abs(js_old - pk.KDE_JS_divergence(a, b, 0.25, 0.0, resolution=256, algo="binned")) <= 1e-12 * js_old
# infinity in one rank's slice raises on every rank
bad = cut(b).copy()
if rank == world - 1:
    bad[3, 1] = np.inf
try:
    D.band_sharded_discrepancy(cut(a), bad, 0.25, 0.0, resolution=100)
    raise SystemExit("infinity was not reported")
except ValueError:
    pass

if os.environ.get("MGPU_SMALL", "0") != "1":
    n = int(os.environ.get("C5_N", 10_000_000))
    A, B = wl.c5_inputs(n, seed=5)
    pool = pk._native.PinnedPool()
    ha = pool.empty(cut(A).shape); ha[:] = cut(A)
    hb = pool.empty(cut(B).shape); hb[:] = cut(B)
    da, db = torch.from_numpy(ha).cuda(), torch.from_numpy(hb).cuda()
    for name, x, y, peer in (("host slices", ha, hb, "1"), ("device slices", da, db, "1"), ("device slices", da, db, "0")):
        os.environ["KDEJS_BAND_PEER"] = peer
        for algo in ("binned", "binned32"):
            ts = []
            for it in range(5):
                info = {}
                torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
                js = D.band_sharded_discrepancy(x, y, 0.25, 0.0, resolution=1024, algo=algo, info=info)
                torch.cuda.synchronize(); dist.barrier(); ts.append(time.perf_counter() - t0)
            if rank == 0:
                print(f"C5 n={n} R=1024 over {world} ranks, {name}, {algo}, exchange over {info['exchange']}: {js!r} in "
                      f"{1e3*min(ts[1:]):.2f} ms", flush=True)
    del os.environ["KDEJS_BAND_PEER"]
    for algo in ("binned", "binned32"):
        info = {"timing": True}
        D.band_sharded_discrepancy(da, db, 0.25, 0.0, resolution=1024, algo=algo, info=info)
        allinfo = [None] * world
        dist.all_gather_object(allinfo, {"rank": rank, "ms": {k: round(v, 3) for k, v in info["timing_ms"].items()},
                                         "cols": info["tile_columns"], "recv": info["points_received"]})
        if rank == 0:
            for x in allinfo:
                print("   steps (synchronised after each)", algo, x, flush=True)
    import ctypes
    nat = pk._native; L = nat.lib()
    def kernel_ms():
        out = {}
        pm, pn = ctypes.c_double(), ctypes.c_int64()
        for k, nm in enumerate(nat.KERNEL_NAMES):
            nat.check(L.kdejs_profile_read(local, k, ctypes.byref(pm), ctypes.byref(pn)))
            if pn.value: out[nm] = round(pm.value, 3)
        return out
    for algo in ("binned", "binned32"):
        nat.check(L.kdejs_profile_reset(local)); nat.check(L.kdejs_profile_enable(local, 1))
        D.band_sharded_discrepancy(da, db, 0.25, 0.0, resolution=1024, algo=algo)
        km = kernel_ms(); nat.check(L.kdejs_profile_enable(local, 0))
        allk = [None] * world
        dist.all_gather_object(allk, km)
        if rank == 0:
            for r_, x in enumerate(allk):
                print(f"   kernel ms (CUDA events) {algo} rank {r_}: {x}", flush=True)
            ta, tb = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
            o = torch.zeros(1, dtype=torch.float64, device="cuda")
            ptr = (ctypes.c_void_p * 1)(tb.data_ptr()); cnt = (ctypes.c_int64 * 1)(len(B))
            st = torch.cuda.Stream()
            call = lambda: nat.check(L.kdejs_discrepancy_batch_dev(local, ta.data_ptr(), len(A), ptr, cnt, 1, 1024, 0.03, 0.25, 0.0, 0, 0, nat.ALGOS[algo], o.data_ptr(), None, st.cuda_stream))
            call(); torch.cuda.synchronize()
            nat.check(L.kdejs_profile_reset(local)); nat.check(L.kdejs_profile_enable(local, 1))
            t0 = time.perf_counter(); call(); torch.cuda.synchronize(); d1 = time.perf_counter() - t0
            print(f"   single GPU device-resident {algo}: {1e3*d1:.2f} ms; kernel ms {kernel_ms()}", flush=True)
            nat.check(L.kdejs_profile_enable(local, 0))
            del ta, tb
        dist.barrier()
    t0 = time.perf_counter()
    js_g = D.row_sharded_discrepancy(ha, hb, 0.25, 0.0, resolution=1024, sliced=True, gather_all=True)
    js_g = D.row_sharded_discrepancy(ha, hb, 0.25, 0.0, resolution=1024, sliced=True, gather_all=True)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    js_g = D.row_sharded_discrepancy(ha, hb, 0.25, 0.0, resolution=1024, sliced=True, gather_all=True)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        print(f"   older variant (all-gather everything, every rank sorts everything): {js_g!r} in {1e3*dt:.2f} ms", flush=True)
        pa = pool.empty(A.shape); pa[:] = A
        pb = pool.empty(B.shape); pb[:] = B
        pk.KDE_JS_divergence(pa, pb, 0.25, 0.0, resolution=1024, algo="binned")
        t0 = time.perf_counter(); full = pk.KDE_JS_divergence(pa, pb, 0.25, 0.0, resolution=1024, algo="binned"); d1 = time.perf_counter() - t0
        print(f"   single GPU from pinned host memory: {full!r} in {1e3*d1:.2f} ms ; rel gap {abs(js_g-full)/full:.2e}", flush=True)
if rank == 0:
    print("mgpu_check ok", flush=True)
dist.barrier()
dist.destroy_process_group()
