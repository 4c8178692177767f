"""GPU probe: one KDE_JS_divergence_batch call over B parameter sets from one pinned arena (config C4), repeated."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poppunk_mod_b200 as pk
from oracle import workloads as wl
nat = pk._native
pool = nat.PinnedPool()
T = pool.empty((100000, 2)); T[:] = wl.example_target()
arena = pool.empty_batch(128, (100000, 2))
for s in range(128): arena[s][:] = wl.sim_batch_member(s)
for B in (512, 2000, 10000, 10000, 10000):
    lst = [arena[k % 128] for k in range(B)]
    t0 = time.perf_counter(); r = pk.KDE_JS_divergence_batch(T, lst, 0.25, 0.0, resolution=256); dt = time.perf_counter() - t0
    print(f"B={B}: {B/dt:.0f} evals/s ({dt*1e3:.1f} ms), {B*1.6e6/dt/1e9:.1f} GB/s", flush=True)
