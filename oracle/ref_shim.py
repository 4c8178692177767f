"""Run the UNMODIFIED reference module -- TEST INFRASTRUCTURE, build container only.

/root/reference/KDE_distance.py imports matplotlib at module top (line 8); matplotlib is
not installed here, and the plotting branch only runs when `outplot` is given.  Two empty
stub modules are registered before the import so the file loads unchanged.  The grid is
hard-coded to 100x100 (KDE_distance.py:89) but `get_grid` is resolved through module
globals, so other resolutions are obtained by rebinding that name -- no edits.

/root/reference does not exist on the GPU box.  oracle/stage_ref.py stages a byte-identical copy of
that one file under oracle/_ref/ (git-ignored, travels with the snapshot); this module loads the
reference tree when it is present and the staged copy otherwise.  Users: oracle/make_golden.py
(fixtures), CPU-side tests, and bench.py's `--impl reference` / cpu_baseline legs.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("KDEJS_REFERENCE_DIR", "/root/reference")
STAGED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_file():
    """Path of the reference's KDE_distance.py: the reference tree, else the staged copy, else None."""
    for d in (REFERENCE_DIR, STAGED_DIR):
        f = os.path.join(d, "KDE_distance.py")
        if os.path.exists(f):
            return f
    return None


def available() -> bool:
    return reference_file() is not None


def load():
    """Import the reference's KDE_distance under the private name `_ref_KDE_distance`."""
    name = "_ref_KDE_distance"
    if name in sys.modules:
        return sys.modules[name]
    if not available():
        raise FileNotFoundError(f"reference not present at {REFERENCE_DIR} nor staged under {STAGED_DIR}")
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules.setdefault("matplotlib", mpl)
        sys.modules.setdefault("matplotlib.pyplot", plt)
    import importlib.util

    spec = importlib.util.spec_from_file_location(name, reference_file())
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    mod._orig_get_grid = mod.get_grid
    return mod


@contextlib.contextmanager
def resolution(R: int):
    """Temporarily make scale_KDE's get_grid(0, 1, 100) call return an R x R grid."""
    mod = load()
    orig = mod._orig_get_grid
    mod.get_grid = lambda mn, mx, _r: orig(mn, mx, R)
    try:
        yield mod
    finally:
        mod.get_grid = orig


def js(df1, df2, gamma=1.0, eps=1e-12, log=False, R: int = 100):
    with resolution(R) as mod:
        return mod.KDE_JS_divergence(df1, df2, gamma, eps, log)


def scale_kde(df1, df2, gamma, eps, R: int = 100):
    """(z1, z2) exactly as the reference's scale_KDE returns them."""
    with resolution(R) as mod:
        z1, z2, _ = mod.scale_KDE(df1, df2, gamma, eps)
    return z1, z2


def raw_density(df1, df2, R: int = 100, log=False, eps=0.0):
    """Reference densities exp(score_samples) of both sets BEFORE **gamma (flat, R*R each),
    obtained by calling the reference's own get_kde / generate_samples on the jointly scaled
    sets (the scaling lines KDE_distance.py:81-86 are replayed with sklearn's MinMaxScaler)."""
    import numpy as np
    from sklearn.preprocessing import MinMaxScaler

    mod = load()
    d1, d2 = np.asarray(df1, dtype=np.float64), np.asarray(df2, dtype=np.float64)
    if log:
        d1, d2 = np.log(d1 + eps), np.log(d2 + eps)
    sc = MinMaxScaler().fit(np.vstack([d1, d2]))
    grid = mod._orig_get_grid(0, 1, R)
    return tuple(mod.generate_samples(grid, mod.get_kde(sc.transform(d))).ravel() for d in (d1, d2))


def js_masked(df1, df2, gamma=1.0, eps=1e-12, log=False, R: int = 100, floor=1e-9):
    """Reference result with its ball-tree log-space residue removed: densities below
    floor*max are set to exactly 0 before **gamma.  Everything else is the
    reference's arithmetic: np.power, += eps, /= sum, scipy jensenshannon."""
    import numpy as np
    from scipy.spatial.distance import jensenshannon

    zs = []
    for d in raw_density(df1, df2, R, log, eps):
        d = d.copy()
        d[d < floor * d.max()] = 0.0
        z = np.power(d, gamma)
        z += eps
        z /= z.sum()
        zs.append(z)
    return jensenshannon(zs[0], zs[1], axis=0), zs[0], zs[1]
