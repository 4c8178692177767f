"""One process, several GPUs: KDE_JS_divergence_batch(devices=[...]) from pageable and pinned arrays
(one host thread and one staging pool per device) must reproduce the single-device result bit for bit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poppunk_mod_b200 as pk
from oracle import workloads as wl
nat = pk._native
ndev = nat.device_count()
obs = wl.synthetic_target(100_000)
sims = [wl.sim_batch_member(s) for s in range(96)]
one = pk.KDE_JS_divergence_batch(obs, sims, 0.25, 0.0, resolution=256, devices=0)
for rep in range(3):
    t0 = time.perf_counter()
    many = pk.KDE_JS_divergence_batch(obs, sims, 0.25, 0.0, resolution=256, devices=list(range(ndev)))
    dt = time.perf_counter() - t0
    assert np.array_equal(one, many), (one[:4], many[:4])
print(f"multidev ok: {ndev} devices, pageable inputs, {len(sims)/dt:.0f} evals/s (last call)")
