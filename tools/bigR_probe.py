import sys, time; sys.path.insert(0,'.')
import numpy as np, poppunk_mod_b200 as pk
from oracle import kde_oracle as ko, workloads as wl
a, b = wl.synthetic_target(20000), wl.sim_S(3, 1500, 0.3, n=15000)
for R in (2048, 4096):
    t0=time.time(); got = pk.KDE_JS_divergence(a, b, 0.25, 0.0, resolution=R); t1=time.time()
    want = ko.discrepancy(a, b, 0.25, 0.0, resolution=R); t2=time.time()
    print(R, got, want, abs(got-want)/want, f"gpu {t1-t0:.3f}s oracle {t2-t1:.1f}s", flush=True)
z1, z2, grid = pk.scale_KDE(a, b, 0.25, 0.0, resolution=2048)
print(z1.shape, z1.sum(), z2.sum())
