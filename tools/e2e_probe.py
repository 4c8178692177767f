import ctypes, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import poppunk_mod_b200 as pk
from oracle import workloads as wl
nat = pk._native; L = nat.lib()
pool = nat.PinnedPool()
T = pool.empty((100000, 2)); T[:] = wl.synthetic_target(100000)
sims = []
for s in range(128):
    b = pool.empty((100000, 2)); b[:] = wl.sim_batch_member(s); sims.append(b)
B = 512
lst = [sims[k % 128] for k in range(B)]
pk.KDE_JS_divergence_batch(T, lst, 0.25, 0.0, resolution=256)
for prof in (0, 1):
    nat.check(L.kdejs_profile_reset(0)); nat.check(L.kdejs_profile_enable(0, prof))
    t0 = time.perf_counter()
    pk.KDE_JS_divergence_batch(T, lst, 0.25, 0.0, resolution=256)
    dt = time.perf_counter() - t0
    print(f"wave={os.environ.get('KDEJS_HOST_WAVE','16')} B={B} prof={prof}: {B/dt:.0f} evals/s ({dt*1e3:.2f} ms)", flush=True)
    if prof:
        tot = 0
        for k, nm in enumerate(nat.KERNEL_NAMES):
            ms, c = ctypes.c_double(), ctypes.c_int64()
            nat.check(L.kdejs_profile_read(0, k, ctypes.byref(ms), ctypes.byref(c)))
            tot += ms.value
            print(f"   {nm:12s} {ms.value:8.3f} ms / {c.value}")
        print(f"   kernels total {tot:.2f} ms")
nat.check(L.kdejs_profile_enable(0, 0))
