import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_grids():
    with np.load(os.path.join(GOLDEN_DIR, "golden_grids.npz")) as f:
        return {k: f[k] for k in f.files}


def golden_inputs(name):
    """Regenerate the inputs of a golden case (same recipes as oracle/make_golden.py:cases)."""
    from oracle import workloads as wl

    a, b = wl.g1_inputs()
    if name.startswith("G1_nan"):
        a = a.copy()
        a[5, 1] = np.nan
        return a, b
    if name.startswith("G1"):
        return a, b
    if name.startswith("G2"):
        return wl.disjoint_inputs()
    if name.startswith("G3"):
        return a, a.copy()
    if name.startswith("G5"):
        return wl.ragged_inputs()
    if name.startswith("G6"):
        return np.column_stack([np.full(500, 0.3), np.linspace(0.1, 0.2, 500)]), b
    if name.startswith("G7"):
        return a[:1], b[:2]
    if name.startswith("G4"):
        return wl.example_target(), wl.sim_S(254, 2500, 0.38)
    if name.startswith("C3_member"):
        return wl.example_target(), wl.sim_batch_member(int(name[len("C3_member"):]))
    raise KeyError(name)


def case_kwargs(case):
    return dict(gamma=case["gamma"], eps=case["eps"], log=case["log"])
