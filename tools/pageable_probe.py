"""Batch call from ordinary (pageable) numpy arrays vs pinned ones: what does a drop-in user get?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poppunk_mod_b200 as pk
from oracle import workloads as wl
nat = pk._native
pool = nat.PinnedPool()
T = wl.synthetic_target(100000)
sims = [wl.sim_batch_member(s) for s in range(128)]
Tp = pool.empty((100000, 2)); Tp[:] = T
simsp = []
for s in sims:
    b = pool.empty((100000, 2)); b[:] = s; simsp.append(b)
for name, t, ss in (("pageable", T, sims), ("pinned", Tp, simsp)):
    for B in (1, 64, 128):
        lst = ss[:B]
        ref = pk.KDE_JS_divergence_batch(t, lst, 0.25, 0.0, resolution=256)
        best = 1e9
        for rep in range(8):
            t0 = time.perf_counter()
            got = pk.KDE_JS_divergence_batch(t, lst, 0.25, 0.0, resolution=256)
            best = min(best, time.perf_counter() - t0)
        print(f"{name:9s} B={B:4d}: {B/best:9.0f} evals/s ({best*1e3:.3f} ms)", flush=True)
