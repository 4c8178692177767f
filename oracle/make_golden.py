"""Generate tests/golden/ by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py

The reference stores no expected values for the KDE discrepancy, so the
fixtures are outputs of /root/reference/KDE_distance.py itself (via oracle/ref_shim.py), with
the library versions recorded.  Three numbers are kept per case:

  ref_raw     KDE_JS_divergence exactly as the reference returns it
  ref_masked  the same with reference densities < 1e-9*max zeroed (removes the ball-tree
              log-space residue that gamma<1 amplifies)
  oracle      oracle/kde_oracle.c (exact fp64 sum)

plus, for a few cases, the per-node reference densities (masked) for the per-grid parity bound.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import kde_oracle as ko  # noqa: E402
from oracle import ref_port, ref_shim, workloads as wl  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
EXAMPLE_TSV = os.path.join(ref_shim.REFERENCE_DIR, "example_input",
                           "core_mu_1e-5_rate_genes1_1e-3_prop_genes2_0.125.tsv")


def cases():
    """name -> (df1, df2, kwargs, R, keep_grids)"""
    T = np.loadtxt(EXAMPLE_TSV, delimiter="\t", dtype="float64")
    a, b = wl.g1_inputs()
    a_nan = a.copy()
    a_nan[5, 1] = np.nan
    da, db = wl.disjoint_inputs()
    ra, rb = wl.ragged_inputs()
    s254 = wl.sim_S(254, 2500, 0.38)
    const = np.column_stack([np.full(500, 0.3), np.linspace(0.1, 0.2, 500)])
    out = {
        "G1_default": (a, b, dict(gamma=1.0, eps=1e-12, log=False), 100, False),
        "G1_g025": (a, b, dict(gamma=0.25, eps=0.0, log=False), 100, True),
        "G1_gridsearch": (a, b, dict(gamma=1e-12, eps=1e-12, log=True), 100, False),
        "G1_nan": (a_nan, b, dict(gamma=0.25, eps=0.0, log=False), 100, False),
        "G1_R256": (a, b, dict(gamma=0.25, eps=0.0, log=False), 256, False),
        "G1_R37": (a, b, dict(gamma=0.25, eps=0.0, log=False), 37, False),
        "G2_disjoint": (da, db, dict(gamma=0.25, eps=0.0, log=False), 100, False),
        "G2_disjoint_default": (da, db, dict(gamma=1.0, eps=1e-12, log=False), 100, False),
        "G3_identical": (a, a.copy(), dict(gamma=0.25, eps=0.0, log=False), 100, False),
        "G5_ragged": (ra, rb, dict(gamma=0.25, eps=0.0, log=False), 100, True),
        "G5_ragged_default": (ra, rb, dict(gamma=1.0, eps=1e-12, log=False), 100, False),
        "G6_const_col": (const, b, dict(gamma=0.25, eps=0.0, log=False), 100, False),
        "G7_tiny": (a[:1], b[:2], dict(gamma=0.25, eps=0.0, log=False), 100, False),
        "G4_R100_g025": (T, s254, dict(gamma=0.25, eps=0.0, log=False), 100, True),
        "G4_R100_g1": (T, s254, dict(gamma=1.0, eps=0.0, log=False), 100, False),
        "G4_R256_g025": (T, s254, dict(gamma=0.25, eps=0.0, log=False), 256, False),
        "G4_gridsearch": (T, s254, dict(gamma=1e-12, eps=1e-12, log=True), 100, False),
    }
    for s in range(4):
        out[f"C3_member{s}"] = (T, wl.sim_batch_member(s), dict(gamma=0.25, eps=0.0, log=False), 100, False)
    return T, out


def main():
    import scipy
    import sklearn

    os.makedirs(GOLDEN, exist_ok=True)
    T, cs = cases()
    np.savez_compressed(os.path.join(GOLDEN, "example_target.npz"), target=T)
    meta = {
        "generated_by": "oracle/make_golden.py",
        "reference": "samhorsfield96/PopPUNK-mod KDE_distance.py (unmodified, imported via oracle/ref_shim.py)",
        "versions": {"numpy": np.__version__, "scipy": scipy.__version__, "scikit-learn": sklearn.__version__,
                     "python": sys.version.split()[0]},
        "reference_pins": {"numpy": "1.26.4", "scipy": "1.11.4", "scikit-learn": "1.7.0"},
        "mask_floor": 1e-9,
        "cases": {},
    }
    grids = {}
    for name, (d1, d2, kw, R, keep) in cs.items():
        t0 = time.time()
        raw = float(ref_shim.js(d1.copy(), d2.copy(), R=R, **kw))
        masked, _, _ = ref_shim.js_masked(d1.copy(), d2.copy(), R=R, **kw)
        port = float(ref_port.js_port(d1.copy(), d2.copy(), resolution=R, **kw))
        orc = float(ko.discrepancy(d1, d2, resolution=R, **kw))
        meta["cases"][name] = dict(kw, R=R, n1=int(len(d1)), n2=int(len(d2)), ref_raw=raw,
                                   ref_masked=float(masked), port=port, oracle=orc)
        if keep:
            r1, r2 = ref_shim.raw_density(d1.copy(), d2.copy(), R=R, log=kw["log"], eps=kw["eps"])
            for tag, r in (("dens1", r1), ("dens2", r2)):
                r = r.copy()
                r[r < 1e-9 * r.max()] = 0.0
                grids[f"{name}_{tag}"] = r
        print(f"{name:24s} raw={raw:.16g} masked={float(masked):.16g} oracle={orc:.16g} "
              f"port==raw:{port == raw} ({time.time() - t0:.1f}s)", flush=True)
    np.savez_compressed(os.path.join(GOLDEN, "golden_grids.npz"), **grids)
    with open(os.path.join(GOLDEN, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
