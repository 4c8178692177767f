"""Config C5 on one GPU: 10 M + 10 M pairs, 1024^2 grid.  Where does the time go?"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poppunk_mod_b200 as pk
from oracle import workloads as wl
nat = pk._native
if os.environ.get("PROBE_LIB"):                       # A/B against another build of the library
    nat.LIB_PATH = os.path.abspath(os.environ["PROBE_LIB"])
L = nat.lib()
n = int(os.environ.get("C5_N", 10_000_000)); R = int(os.environ.get("C5_R", 1024))
a = wl.synthetic_target(n, seed=11); b = wl.sim_S(5, 1800, 0.36, n=n)
pool = nat.PinnedPool(); pa = pool.empty((n, 2)); pa[:] = a; pb = pool.empty((n, 2)); pb[:] = b
for name, x, y in (("pageable", a, b), ("pinned", pa, pb)):
    pk.KDE_JS_divergence(x, y, 0.25, 0.0, resolution=R)
    nat.check(L.kdejs_profile_reset(0)); nat.check(L.kdejs_profile_enable(0, 1))
    t0 = time.perf_counter(); js = pk.KDE_JS_divergence(x, y, 0.25, 0.0, resolution=R); dt = time.perf_counter() - t0
    print(f"{name}: js={js!r} wall {dt*1e3:.1f} ms")
    for k, nm in enumerate(nat.KERNEL_NAMES):
        ms, c = ctypes.c_double(), ctypes.c_int64()
        nat.check(L.kdejs_profile_read(0, k, ctypes.byref(ms), ctypes.byref(c)))
        print(f"   {nm:12s} {ms.value:8.3f} ms / {c.value}")
    nat.check(L.kdejs_profile_enable(0, 0))
