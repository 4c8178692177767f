"""Dense kernels with the Gaussian kernel: timing and agreement (fp32 separable vs fp64 brute force)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poppunk_mod_b200 as pk
from oracle import workloads as wl
n = int(os.environ.get("GP_N", 100_000)); R = int(os.environ.get("GP_R", 256))
a = wl.synthetic_target(n, seed=11); b = wl.sim_S(5, 1800, 0.36, n=n)
for algo in ("dense", "dense64"):
    for kern in ("epanechnikov", "gaussian"):
        pk.KDE_JS_divergence(a, b, 0.25, 0.0, resolution=R, algo=algo, kernel=kern)
        t0 = time.perf_counter()
        js = pk.KDE_JS_divergence(a, b, 0.25, 0.0, resolution=R, algo=algo, kernel=kern)
        dt = time.perf_counter() - t0
        print(f"{algo:8s} {kern:13s} js={js!r} {dt*1e3:9.2f} ms  {2*n*R*R/dt/1e12:6.2f} Tpairs/s")
