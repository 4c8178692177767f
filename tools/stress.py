"""Randomised parity stress on a GPU: many random configurations against the CPU oracle (test infrastructure)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poppunk_mod_b200 as pk
from oracle import kde_oracle as ko

# STRESS_ALGO=binned (fp64 sum: 1e-9 against the oracle) or binned32 (the default algorithm: the north star's 1e-5)
ALGO = os.environ.get("STRESS_ALGO", "binned")
TOL = 1e-9 if ALGO == "binned" else 1e-5
n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 400
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
worst, nan_cases, t0 = 0.0, 0, time.time()
batch_obs, batch_sims, batch_want = None, [], []
for trial in range(n_trials):
    n1, n2 = int(rng.integers(1, 30000)), int(rng.integers(1, 30000))
    R = int(rng.choice([3, 9, 33, 64, 100, 100, 127, 256, 300]))
    h = float(rng.choice([0.004, 0.01, 0.03, 0.03, 0.03, 0.1, 0.5]))
    gamma = float(rng.choice([0.25, 0.25, 0.5, 1.0, 0.05, 1e-12, 3.0]))
    eps = float(rng.choice([0.0, 0.0, 1e-12, 1e-5]))
    log = bool(rng.integers(0, 5) == 0)
    shape = rng.choice(["blob", "curve", "lattice", "line"])
    def cloud(n):
        if shape == "blob":
            return np.abs(rng.normal([0.002, 0.3], [rng.uniform(1e-5, 1e-3), rng.uniform(1e-3, 0.2)], (n, 2))) + 1e-9
        if shape == "curve":
            x = rng.beta(2, 2, n) * 0.0016
            return np.column_stack([x + 1e-9, np.clip(0.38 - 0.26 * np.exp(-rng.uniform(500, 5000) * x) + rng.normal(0, 0.01, n), 1e-9, 1)])
        if shape == "lattice":
            return np.column_stack([rng.integers(1, 40, n) / 1e4, rng.integers(1, 60, n) / 100.0])
        return np.column_stack([np.full(n, 0.001), rng.random(n) + 1e-9])
    a, b = cloud(n1), cloud(n2)
    if rng.integers(0, 4) == 0:
        b[rng.integers(0, n2, max(1, n2 // 50)), rng.integers(0, 2)] = np.nan
    got = pk.KDE_JS_divergence(a, b, gamma, eps, log, resolution=R, bandwidth=h, algo=ALGO)
    want = ko.discrepancy(a, b, gamma, eps, log, resolution=R, bandwidth=h)
    if np.isnan(want):      # 0/0 (no node in reach) -> NaN on both sides; or sqrt of -1e-17 noise in the oracle's
        assert np.isnan(got) or got < 3e-8, (trial, got, want)      # left/right sums, where the product returns ~0
        nan_cases += 1
        continue
    # JS^2 is a difference of O(1e-3) sums in the reference's (and the oracle's) arithmetic: below JS ~ 1e-7 the
    # value is rounding noise there, so tiny distances are compared absolutely
    err = abs(got - want) / max(abs(want), 1e-12)
    if abs(got - want) > 3e-8:
        worst = max(worst, err)
        assert err < TOL, (trial, n1, n2, R, h, gamma, eps, log, shape, got, want)
    elif want > 1e-6:
        worst = max(worst, err)
        assert err < TOL, (trial, n1, n2, R, h, gamma, eps, log, shape, got, want)
print(f"stress ok ({ALGO}, bound {TOL:g}): {n_trials} random configurations, worst relative error {worst:.2e}, NaN-consistent cases {nan_cases}, "
      f"{time.time() - t0:.0f} s")

# ---- batches equal singles bitwise; dense fp32 kernel within its tolerance ---------------------------------
t0 = time.time()
for trial in range(12):
    nobs = int(rng.integers(100, 20000))
    obs = np.abs(rng.normal([0.002, 0.3], [3e-4, 0.05], (nobs, 2)))
    sims = [np.abs(rng.normal([0.002, 0.3], [rng.uniform(1e-4, 6e-4), rng.uniform(0.01, 0.1)],
                              (int(rng.integers(1, 15000)), 2))) for _ in range(int(rng.integers(1, 70)))]
    R = int(rng.choice([50, 100, 128]))
    out = pk.KDE_JS_divergence_batch(obs, sims, 0.25, 0.0, resolution=R, algo=ALGO)
    for i in rng.choice(len(sims), size=min(4, len(sims)), replace=False):
        single = pk.KDE_JS_divergence(obs, sims[i], 0.25, 0.0, resolution=R, algo=ALGO)
        assert out[i] == single, (trial, i, out[i], single)
        dense = pk.KDE_JS_divergence(obs, sims[i], 0.25, 0.0, resolution=R, algo="dense")
        assert abs(dense - single) <= 1e-5 * single + 1e-7, (trial, i, dense, single)
print(f"batch == singles bitwise and dense fp32 within 1e-5 on 12 random batches, {time.time() - t0:.0f} s")
