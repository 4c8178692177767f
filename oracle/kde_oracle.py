"""ctypes front-end of the CPU oracle (oracle/kde_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.

The oracle restates KDE_distance.KDE_JS_divergence (/root/reference/KDE_distance.py:104-130)
as an exact fp64 sum; see the header of kde_oracle.c for the line-by-line map.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libkdeoracle.so")
_lib = None

KERNELS = {"epanechnikov": 0, "gaussian": 1}
SQRT_LN2 = 0.8325546111576977  # upper bound of the JS distance (natural log)


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc, a second or two)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "kde_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.oracle_discrepancy.restype = ctypes.c_int
        lib.oracle_discrepancy.argtypes = [
            dp, ctypes.c_int64, dp, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
            ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            dp, dp, dp, dp, dp]
        lib.oracle_fit_scaler.restype = ctypes.c_int
        lib.oracle_fit_scaler.argtypes = [dp, ctypes.c_int64, dp, ctypes.c_int64,
                                          ctypes.c_int, ctypes.c_double, dp, dp]
        lib.oracle_transform.restype = None
        lib.oracle_transform.argtypes = [dp, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
                                         dp, dp, dp]
        for name in ("oracle_kde_dense",):
            f = getattr(lib, name)
            f.restype = None
            f.argtypes = [dp, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                          ctypes.c_double, ctypes.c_int, dp]
        lib.oracle_kde_window.restype = None
        lib.oracle_kde_window.argtypes = [dp, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                          ctypes.c_int, ctypes.c_double, dp]
        lib.oracle_pow_norm.restype = None
        lib.oracle_pow_norm.argtypes = [dp, ctypes.c_int64, ctypes.c_double, ctypes.c_double, dp]
        lib.oracle_js.restype = ctypes.c_double
        lib.oracle_js.argtypes = [dp, dp, ctypes.c_int64]
        _lib = lib
    return _lib


def _as_pts(x) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if a.ndim != 2 or a.shape[1] != 2:
        raise ValueError(f"expected an (N, 2) array, got shape {a.shape}")
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def discrepancy(df1, df2, gamma=1.0, eps=1e-12, log=False, *, resolution=100, bandwidth=0.03,
                kernel="epanechnikov", dense=False, return_grids=False):
    """Exact-sum JS distance; with return_grids also (z1, z2, dens1, dens2), each (R*R,)."""
    a, b = _as_pts(df1), _as_pts(df2)
    lib = _load()
    G = resolution * resolution
    out = ctypes.c_double()
    bufs = [np.empty(G) for _ in range(4)] if return_grids else [None] * 4
    rc = lib.oracle_discrepancy(_p(a), a.shape[0], _p(b), b.shape[0], resolution, bandwidth,
                                gamma, eps, int(bool(log)), KERNELS[kernel], int(bool(dense)),
                                ctypes.byref(out), *[_p(x) for x in bufs])
    if rc == 1:
        raise ValueError("Input contains infinity or a value too large for dtype('float64').")
    if rc != 0:
        raise ValueError("oracle: bad arguments")
    js = np.float64(out.value)
    return (js, *bufs) if return_grids else js


def fit_scaler(df1, df2, log=False, eps=0.0):
    a, b = _as_pts(df1), _as_pts(df2)
    scale, min_ = np.empty(2), np.empty(2)
    has_inf = _load().oracle_fit_scaler(_p(a), a.shape[0], _p(b), b.shape[0], int(bool(log)), eps,
                                        _p(scale), _p(min_))
    return scale, min_, bool(has_inf)


def transform(df, scale, min_, log=False, eps=0.0):
    a = _as_pts(df)
    out = np.empty_like(a)
    _load().oracle_transform(_p(a), a.shape[0], int(bool(log)), eps,
                             _p(np.ascontiguousarray(scale)), _p(np.ascontiguousarray(min_)), _p(out))
    return out


def kde_density(X_scaled, resolution=100, bandwidth=0.03, kernel="epanechnikov", dense=False,
                gmin=0.0, gmax=1.0):
    """exp(score_samples(get_grid(gmin, gmax, R)[2])) as an exact sum, flat (R*R,)."""
    X = _as_pts(X_scaled)
    out = np.empty(resolution * resolution)
    lib = _load()
    if dense or kernel != "epanechnikov":
        lib.oracle_kde_dense(_p(X), X.shape[0], gmin, gmax, resolution, bandwidth, KERNELS[kernel], _p(out))
    else:
        lib.oracle_kde_window(_p(X), X.shape[0], gmin, gmax, resolution, bandwidth, _p(out))
    return out


def pow_norm(dens, gamma, eps):
    d = np.ascontiguousarray(dens, dtype=np.float64).ravel()
    z = np.empty_like(d)
    _load().oracle_pow_norm(_p(d), d.size, gamma, eps, _p(z))
    return z


def js_distance(p, q):
    p = np.ascontiguousarray(p, dtype=np.float64).ravel()
    q = np.ascontiguousarray(q, dtype=np.float64).ravel()
    return np.float64(_load().oracle_js(_p(p), _p(q), p.size))
