"""Small run through every kernel path, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poppunk_mod_b200 as pk
from oracle import kde_oracle as ko, workloads as wl

a, b = wl.ragged_inputs()
for R in (37, 100):
    for algo in ("binned", "dense", "dense64"):
        got = pk.KDE_JS_divergence(a, b, 0.25, 0.0, resolution=R, algo=algo)
        want = ko.discrepancy(a, b, 0.25, 0.0, resolution=R)
        assert abs(got - want) <= 1e-5 * want, (R, algo, got, want)
obs = wl.synthetic_target(6000)
sims = [wl.sim_batch_member(s, n=1500 + 211 * s) for s in range(9)]
out = pk.KDE_JS_divergence_batch(obs, sims, 0.25, 0.0)
assert abs(out[3] - ko.discrepancy(obs, sims[3], 0.25, 0.0)) < 1e-10
pk.KDE_JS_divergence_pairs(sims[:4], [(0, 1), (2, 3), (1, 3)], 0.25, 0.0)
pk.kde_density(np.clip(sims[0], 0, 1), resolution=50)
pk.KDE_JS_divergence(a[:300], b[:200], 1.0, 1e-12, kernel="gaussian", resolution=20, algo="dense")
pk.KDE_JS_divergence(a, b, 1e-12, 1e-12, True)
print("sanitize driver ok")
