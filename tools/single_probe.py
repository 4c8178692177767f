"""GPU probe: ONE evaluation per call (the BOLFI acquisition phase): latency from host memory and per-kernel times."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import poppunk_mod_b200 as pk
from oracle import workloads as wl
nat = pk._native; L = nat.lib(); dev = 0
pool = nat.PinnedPool()
T = pool.empty((100000, 2)); T[:] = wl.example_target()
S = pool.empty((100000, 2)); S[:] = wl.sim_batch_member(3)
for R in (100, 256):
    for algo in ("binned32", "binned"):
        for _ in range(5): pk.KDE_JS_divergence(T, S, 0.25, 0.0, resolution=R, algo=algo)
        t0 = time.perf_counter()
        for _ in range(200): js = pk.KDE_JS_divergence(T, S, 0.25, 0.0, resolution=R, algo=algo)
        dt = (time.perf_counter() - t0) / 200
        nat.check(L.kdejs_profile_reset(dev)); nat.check(L.kdejs_profile_enable(dev, 1))
        for _ in range(20): pk.KDE_JS_divergence(T, S, 0.25, 0.0, resolution=R, algo=algo)
        pm, pn = ctypes.c_double(), ctypes.c_int64(); prof = {}
        for k, nm in enumerate(nat.KERNEL_NAMES):
            nat.check(L.kdejs_profile_read(dev, k, ctypes.byref(pm), ctypes.byref(pn)))
            if pn.value: prof[nm] = round(1e3 * pm.value / 20, 1)
        nat.check(L.kdejs_profile_enable(dev, 0))
        print(f"R={R} {algo}: {1e3*dt:.3f} ms per call from pinned host memory ({1/dt:.0f}/s); kernel us {prof} sum {sum(prof.values()):.0f}", flush=True)
