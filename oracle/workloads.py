"""Seeded synthetic inputs shared by tests, bench.py and oracle/make_golden.py.

Numpy only, no reference access: everything here can be regenerated on the GPU box.
Shapes follow SURVEY.md section 8(d): (core, accessory) distance pairs as Pansim writes them
(/root/reference/poppunk_mod.py:360, two tab-separated float columns).
"""
from __future__ import annotations

import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def sim_S(seed, b2=2500.0, a0=0.38, n=100_000):
    """Simulated cloud: core ~ 0.0016*Beta(2,2); accessory follows the negative-exponential
    curve of poppunk_mod.py:39-40 plus N(0, 0.01) noise, clipped to [0,1]."""
    r = np.random.default_rng(seed)
    x = r.beta(2, 2, n) * 0.0016
    y = a0 - (a0 - 0.12) * np.exp(-b2 * x) + r.normal(0, 0.01, n)
    return np.column_stack([x, np.clip(y, 0, 1)])


def sim_batch_params(seed):
    """(b2, a0) for batch member `seed` (C3/C4 of SURVEY.md 8d)."""
    r = np.random.default_rng(10_000_019 + seed)
    return float(r.uniform(500, 5000)), float(r.uniform(0.2, 0.5))


def sim_batch_member(seed, n=100_000):
    b2, a0 = sim_batch_params(seed)
    return sim_S(seed, b2, a0, n)


def synthetic_target(n=100_000, seed=7):
    """Stand-in for example_input/core_mu_1e-5_...tsv where the real file is not needed:
    discrete core values (Pansim core distances are k/L for integer k), accessory = ratio noise."""
    r = np.random.default_rng(seed)
    k = r.binomial(1_000_000, 0.0009 * r.beta(6, 3, n))
    x = k / 1_000_000.0
    y = 0.38 - 0.26 * np.exp(-1800.0 * x) + r.normal(0, 0.012, n)
    return np.column_stack([x, np.clip(y, 0, 1)])


def example_target():
    """The reference's example target (100 000 x 2), stored losslessly as a binary fixture
    (tests/golden/example_target.npz, written by oracle/make_golden.py)."""
    with np.load(os.path.join(_GOLDEN, "example_target.npz")) as f:
        return f["target"].astype(np.float64)


def c5_inputs(n=10_000_000, seed=5):
    """Config C5 of SURVEY.md 8(d): a target of n real-genome-shaped pairs, bootstrap-resampled with jitter
    N(0, (1e-4 * range)^2) from real rows (the reference's example target -- the committed fixture that travels
    to the GPU box; distances/2_column is only in the build container), and a simulated set of n pairs from the
    S generator stretched to the target's range."""
    base = example_target()
    r = np.random.default_rng(seed)
    span = base.max(0) - base.min(0)
    idx = r.integers(0, len(base), n)
    target = base[idx] + r.standard_normal((n, 2)) * (1e-4 * span)
    sim = sim_S(seed + 1, 1800.0, 0.36, n=n)
    lo, hi = sim.min(0), sim.max(0)
    sim = (sim - lo) / (hi - lo) * span + base.min(0)
    return np.ascontiguousarray(target), np.ascontiguousarray(sim)


def g1_inputs():
    """SURVEY.md 8(c) golden case G1 (same generator, drawn in this order)."""
    rng = np.random.default_rng(0)
    a = rng.random((2000, 2)) * [0.01, 0.4]
    b = rng.random((3000, 2)) * [0.02, 0.3] + [0, 0.05]
    return a, b


def disjoint_inputs():
    rng = np.random.default_rng(1)
    a = rng.random((1500, 2)) * [0.2, 0.2]
    b = rng.random((1500, 2)) * [0.2, 0.2] + [0.8, 0.8]
    return a, b


def ragged_inputs(seed=3):
    """Unequal sizes, duplicated rows, a constant column in one set, NaNs in both."""
    rng = np.random.default_rng(seed)
    a = rng.normal([0.002, 0.3], [0.0005, 0.05], (777, 2))
    b = rng.normal([0.003, 0.25], [0.001, 0.08], (4099, 2))
    b[:, 0] = np.round(b[:, 0], 4)           # heavy ties, as in Pansim output
    a[::50, 0] = np.nan
    b[7, 1] = np.nan
    b[100:140] = b[100]
    return a, b
