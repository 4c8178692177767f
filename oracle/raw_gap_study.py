"""How far is the UNMODIFIED reference from the exact sum where its callers use it?  TEST INFRASTRUCTURE, build
container only (needs /root/reference or the staged copy).

The reference's ball tree leaves ~1e-15 residues at grid nodes whose true density is 0; `z**gamma` with gamma < 1
lifts them.  This script scores N synthetic simulations against the reference's example target at
the reference's own grid (100 x 100) and its two production call sites

    BOLFI / pairwise:  KDE_JS_divergence(obs, sim, gamma=0.25, eps=0, log=False)       poppunk_mod.py:373,
                                                                                       scripts/pairwise_2D_distance.py:54
    grid search:       KDE_JS_divergence(ref, sim, 1e-12, 1e-12, True)                 scripts/distance_to_gridsearch.py:50

with (a) the unmodified reference and (b) the exact oracle (what the GPU path returns to ~1e-13), and writes
tests/golden/raw_gap_study.json: per-simulation values plus what the gap does to the DECISIONS downstream --
the arg-min the grid search keeps (distance_to_gridsearch.py:111-125), the ranking, and BOLFI's log-discrepancy
(poppunk_mod.py:571-572; with matching gene-frequency components d is the JS component itself).

    python -m oracle.raw_gap_study [N=200] [processes=6]
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(_HERE, "..", "tests", "golden", "raw_gap_study.json")
SITES = {"bolfi_g025": dict(gamma=0.25, eps=0.0, log=False),
         "gridsearch_g1e-12_log": dict(gamma=1e-12, eps=1e-12, log=True)}


def _job(seed):
    from oracle import kde_oracle, ref_shim
    from oracle import workloads as wl

    T = wl.example_target()
    S = wl.sim_batch_member(seed)
    row = {"seed": seed}
    for name, kw in SITES.items():
        row[name] = {"ref_raw": float(ref_shim.js(T.copy(), S.copy(), kw["gamma"], kw["eps"], kw["log"], R=100)),
                     "exact": float(kde_oracle.discrepancy(T, S, kw["gamma"], kw["eps"], kw["log"], resolution=100))}
    return row


def spearman(x, y):
    rx, ry = np.argsort(np.argsort(x)), np.argsort(np.argsort(y))
    return float(np.corrcoef(rx, ry)[0, 1])


def summarise(rows):
    out = {}
    for name in SITES:
        raw = np.array([r[name]["ref_raw"] for r in rows])
        ex = np.array([r[name]["exact"] for r in rows])
        relgap = np.abs(raw - ex) / ex
        dlog = np.log(raw) - np.log(ex)
        order_raw, order_ex = np.argsort(raw), np.argsort(ex)
        # how much worse (in the reference's own metric) is the file the exact values would keep?
        regret = float(raw[order_ex[0]] - raw[order_raw[0]])
        out[name] = {
            "n": len(rows), "rel_gap_max": float(relgap.max()), "rel_gap_mean": float(relgap.mean()),
            "rel_gap_min": float(relgap.min()),
            "log_d_shift_max": float(np.abs(dlog).max()), "log_d_shift_mean": float(dlog.mean()),
            "log_d_shift_std": float(dlog.std()),
            "argmin_seed_ref_raw": int(rows[int(order_raw[0])]["seed"]), "argmin_seed_exact": int(rows[int(order_ex[0])]["seed"]),
            "argmin_agrees": bool(order_raw[0] == order_ex[0]),
            "argmin_regret_in_reference_metric": regret,
            "top10_overlap": int(len(set(order_raw[:10]) & set(order_ex[:10]))),
            "spearman_rho": spearman(raw, ex),
            "pairwise_order_flips": int(sum((raw[i] < raw[j]) != (ex[i] < ex[j])
                                            for i in range(len(raw)) for j in range(i + 1, len(raw)))),
            "pairs": len(raw) * (len(raw) - 1) // 2,
        }
    return out


def main(n=200, processes=6):
    import multiprocessing as mp

    import scipy
    import sklearn

    with mp.get_context("fork").Pool(processes) as pool:
        rows = pool.map(_job, list(range(n)))
    doc = {"generator": "oracle/raw_gap_study.py", "resolution": 100, "target": "example target (100 000 real rows)",
           "sims": "oracle.workloads.sim_batch_member(seed), seed = 0..n-1",
           "versions": {"numpy": np.__version__, "scipy": scipy.__version__, "sklearn": sklearn.__version__},
           "summary": summarise(rows), "rows": rows}
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc["summary"], indent=1))


if __name__ == "__main__":
    main(*(int(x) for x in sys.argv[1:3]))
