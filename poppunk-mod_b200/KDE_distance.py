"""Drop-in for PopPUNK-mod's ``KDE_distance`` module, computed on a B200 through libkdejs.so.

Same public names, argument meaning and return types as /root/reference/KDE_distance.py:

    get_grid(minimum, maximum, resolution)                       KDE_distance.py:17-23
    scale_KDE(df1, df2, gamma, eps) -> (z1, z2, grid_params)     KDE_distance.py:80-95
    KDE_JS_divergence(df1, df2, gamma=1.0, eps=1e-12,
                      log=False, outplot=None) -> np.float64     KDE_distance.py:104-130
    KDE_KL_divergence(df1, df2, gamma=1.0, eps=1e-12)            KDE_distance.py:97-102
    main()  (python KDE_distance.py --infile1 A --infile2 B)     KDE_distance.py:132-144

Callers that keep working unchanged: poppunk_mod.py:373, scripts/pairwise_2D_distance.py:54,
scripts/distance_to_gridsearch.py:50.  The reference's hard-coded 100x100 grid, bandwidth 0.03
and Epanechnikov kernel (KDE_distance.py:89, :50-51) are the defaults of extra keyword-only
arguments.  Added entry points: KDE_JS_divergence_batch / _pairs (many evaluations per call,
optionally over several GPUs) and kde_density.

All arithmetic of the path runs in hand-written CUDA kernels (poppunk-mod_b200/csrc); there is
no CPU fallback -- without the library or a B200 the calls raise RuntimeError.

Known, deliberate difference: the reference's sklearn ball tree leaves ~1e-15 residues at grid
nodes whose density is exactly zero; with gamma < 1 those residues move its result by ~5e-4
(SURVEY.md section 0.3).  This module returns the exact sum, which equals the reference with
those residues removed to ~1e-13 (tests/test_parity_gpu.py, tests/golden/golden.json).
"""
from __future__ import annotations

import argparse
import ctypes
import os
import sys
import threading

import numpy as np

try:
    from . import _native as nat
except ImportError:  # used as a top-level drop-in module: `sys.path.insert(0, ".../poppunk-mod_b200")`
    import importlib.util as _ilu

    _spec = _ilu.spec_from_file_location("_kdejs_native",
                                         os.path.join(os.path.dirname(os.path.abspath(__file__)), "_native.py"))
    nat = _ilu.module_from_spec(_spec)
    _spec.loader.exec_module(nat)

__all__ = ["get_grid", "scale_KDE", "KDE_JS_divergence", "KDE_KL_divergence", "KDE_JS_divergence_batch",
           "KDE_JS_divergence_batch_text", "KDE_JS_divergence_pairs", "kde_density", "summary_batch", "summary_from_js", "rule_bandwidth", "default_device", "main"]

_DEFAULT_RESOLUTION = 100      # KDE_distance.py:89
_DEFAULT_BANDWIDTH = 0.03      # KDE_distance.py:50
# Density algorithm used when a caller does not name one (Epanechnikov kernel):
#   "binned32"  the culled sum with fp32 terms on the FP32 pipe: exact zeros stay exact (nodes the fp32 arithmetic
#               cannot certify are redone in fp64), non-zero node densities within ~1e-7 relative, JS within ~1e-8
#               of the exact sum (bar: 1e-5); geometries its fixed-point format does not cover, sets of fewer than
#               64 points and fine grids with dense cells (moment cells) run the fp64 kernel anyway;
#   "binned"    the same sum entirely in fp64 (the reference's exact arithmetic to ~1e-16), ~20 % slower.
# KDEJS_DEFAULT_ALGO overrides the choice for a whole process.
DEFAULT_ALGO = os.environ.get("KDEJS_DEFAULT_ALGO", "binned32")


def default_device() -> int:
    """KDEJS_DEVICE, else LOCAL_RANK (one process per GPU under torchrun), else 0."""
    for key in ("KDEJS_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(key)
        if v not in (None, ""):
            n = nat.device_count()
            return int(v) % n if n > 0 else int(v)
    return 0


def get_grid(minimum, maximum, resolution):
    """(xx, yy, xy) with xy[k] = (lin[k // R], lin[k % R]); KDE_distance.py:17-23."""
    lin = np.linspace(minimum, maximum, resolution)
    xx, yy = np.meshgrid(lin, lin)
    xy = np.column_stack([yy.reshape(-1), xx.reshape(-1)])
    return xx, yy, xy


def rule_bandwidth(rule, n_samples, n_features=2):
    """sklearn's bandwidth rules (sklearn:neighbors/_kde.py:223-229): "scott" = n**(-1/(d+4)),
    "silverman" = (n (d+2) / 4)**(-1/(d+4)); no covariance factor.  For d = 2 the two coincide."""
    if rule == "scott":
        return float(n_samples) ** (-1.0 / (n_features + 4))
    if rule == "silverman":
        return (float(n_samples) * (n_features + 2) / 4.0) ** (-1.0 / (n_features + 4))
    raise ValueError(f"bandwidth must be a positive number, 'scott' or 'silverman'; got {rule!r}")


def _bandwidth(bandwidth, *sets):
    """A number passes through; a rule name is resolved from the sample count like sklearn does at fit() time.
    The kernels use one bandwidth per evaluation, so with a rule all sets of the call must have the same size."""
    if not isinstance(bandwidth, str):
        return float(bandwidth)
    sizes = set()
    for x in sets:
        sizes |= x.sizes() if isinstance(x, nat.PointSets) else {int(np.shape(x)[0])}
    if len(sizes) != 1:
        raise ValueError(f"bandwidth={bandwidth!r} needs point sets of equal size in one call (sklearn fits the rule "
                         f"per set; sizes here: {sorted(sizes)}); pass a number instead")
    return rule_bandwidth(bandwidth, sizes.pop())


def _common(resolution, bandwidth, gamma, eps, log, kernel, algo):
    if kernel not in nat.KERNELS:
        raise ValueError(f"kernel must be one of {sorted(nat.KERNELS)}; got {kernel!r}")
    if algo is None:
        algo = DEFAULT_ALGO if kernel == "epanechnikov" else "dense64"
    if algo not in nat.ALGOS:
        raise ValueError(f"algo must be one of {sorted(nat.ALGOS)}; got {algo!r}")
    return (int(resolution), float(bandwidth), float(gamma), float(eps), int(bool(log)),
            nat.KERNELS[kernel], nat.ALGOS[algo])


def _discrepancy(df1, df2, gamma, eps, log, resolution, bandwidth, kernel, algo, device, want_z=False,
                 want_density=False):
    a = nat.as_points(df1, "df1")
    b = nat.as_points(df2, "df2")
    common = _common(resolution, _bandwidth(bandwidth, a, b), gamma, eps, log, kernel, algo)
    G = int(resolution) * int(resolution)
    out = ctypes.c_double()
    z1 = z2 = d1 = d2 = None
    if want_z:
        z1, z2 = np.empty(G), np.empty(G)
    if want_density:
        d1, d2 = np.empty(G), np.empty(G)
    dev = default_device() if device is None else int(device)
    nat.check(nat.lib().kdejs_discrepancy(dev, nat.ptr(a), a.shape[0], nat.ptr(b), b.shape[0], *common,
                                          ctypes.byref(out), nat.ptr(z1), nat.ptr(z2), nat.ptr(d1), nat.ptr(d2)))
    return np.float64(out.value), z1, z2, d1, d2


def scale_KDE(df1, df2, gamma, eps, *, log=False, resolution=_DEFAULT_RESOLUTION,
              bandwidth=_DEFAULT_BANDWIDTH, kernel="epanechnikov", algo=None, device=None):
    """Jointly min-max scale both sets, KDE each on the grid, **gamma, +eps, normalise."""
    _, z1, z2, _, _ = _discrepancy(df1, df2, gamma, eps, log, resolution, bandwidth, kernel, algo, device,
                                   want_z=True)
    return z1, z2, get_grid(0, 1, resolution)


def KDE_KL_divergence(df1, df2, gamma=1.0, eps=1e-12, **kw):
    """sum(rel_entr(z1, z2)).  (The reference's version unpacks two of scale_KDE's three return
    values and raises ValueError, KDE_distance.py:98; this one returns the intended number.)"""
    z1, z2, _ = scale_KDE(df1, df2, gamma, eps, **kw)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(z1 > 0, z1 * np.log(z1 / z2), np.where(z1 == 0, 0.0, np.inf))
    return np.float64(t.sum())


def KDE_JS_divergence(df1, df2, gamma=1.0, eps=1e-12, log=False, outplot=None, *,
                      resolution=_DEFAULT_RESOLUTION, bandwidth=_DEFAULT_BANDWIDTH,
                      kernel="epanechnikov", algo=None, device=None):
    """Jensen-Shannon distance between the grid KDEs of two (N,2) point sets."""
    want_plot = outplot is not None
    js, z1, z2, _, _ = _discrepancy(df1, df2, gamma, eps, log, resolution, bandwidth, kernel, algo, device,
                                    want_z=want_plot)
    if want_plot:
        _plot_contours(df1, df2, z1, z2, eps, log, resolution, outplot)
    return js


def KDE_JS_divergence_batch(obs, sims, gamma=1.0, eps=1e-12, log=False, *,
                            resolution=_DEFAULT_RESOLUTION, bandwidth=_DEFAULT_BANDWIDTH,
                            kernel="epanechnikov", algo=None, devices=None):
    """[KDE_JS_divergence(obs, s, gamma, eps, log) for s in sims] as one float64 array.

    sims: a sequence of (N_i,2) arrays, or one (B,N,2) array (no per-simulation Python work then).
    devices: None (default device), an int, or a list of device indices; with several devices the
    sims are dealt round-robin, one host thread per device, and only the B scalars come back."""
    o = nat.as_points(obs, "obs")
    sets = nat.PointSets(sims, "sims")
    common = _common(resolution, _bandwidth(bandwidth, o, sets), gamma, eps, log, kernel, algo)
    B = len(sets)
    out = np.empty(B, dtype=np.float64)
    if B == 0:
        return out
    if devices is None:
        devices = [default_device()]
    elif isinstance(devices, int):
        devices = [devices]
    devices = list(devices)

    def run(dev, idx):
        ptrs, counts, n, keep = sets.pointers(idx)
        res = np.empty(n, dtype=np.float64)
        rc = nat.lib().kdejs_discrepancy_batch(int(dev), nat.ptr(o), o.shape[0], ptrs, counts, n,
                                               *common, nat.ptr(res))
        err = (nat.lib().kdejs_last_error() or b"").decode() if rc else ""
        return idx, res, rc, err

    if len(devices) == 1:
        results = [run(devices[0], slice(None))]
    else:
        results = [None] * len(devices)
        shards = [slice(k, B, len(devices)) for k in range(len(devices))]

        def worker(k):
            results[k] = run(devices[k], shards[k]) if k < B else (slice(0, 0), np.empty(0), 0, "")

        threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(devices))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    first_err = None
    for idx, res, rc, err in results:
        out[idx] = res
        if rc and first_err is None:
            first_err = (rc, err)
    if first_err:
        rc, err = first_err
        if rc in (nat.KDEJS_ERR_ARG, nat.KDEJS_ERR_INF):
            e = ValueError(err)
            if rc == nat.KDEJS_ERR_INF:
                e.partial_results = out      # the other evaluations finished; the offending ones are NaN
            raise e
        raise RuntimeError(f"kdejs: {err}")
    return out


def KDE_JS_divergence_batch_text(obs, texts, gamma=1.0, eps=1e-12, log=False, *, nbytes=None, skip_rows=-1,
                                 resolution=_DEFAULT_RESOLUTION, bandwidth=_DEFAULT_BANDWIDTH, kernel="epanechnikov",
                                 algo=None, device=None, bounds=False):
    """[KDE_JS_divergence(obs, table_i)] for tables given as the BYTES of their files: the text is uploaded as it is
    and parsed on the device (last two tab-separated columns, `skip_rows` as in TsvTable; kdejs_discrepancy_batch_text).

    texts: a sequence of bytes-like objects / uint8 arrays, or ONE (B, slot_bytes) uint8 array with `nbytes[i]` valid
    bytes in row i (a _native.TextArena).  Returns (js, rows, status[, bounds]): status[i] is _native.KDEJS_TABLE_OK,
    KDEJS_TABLE_HOST (js[i] is NaN: the table holds something only the host reader covers -- a missing-value spelling,
    inf / nan, more than 19 digits, a malformed row -- or no rows) or KDEJS_TABLE_INF; bounds[i] = (min, max) of
    column 0, (min, max) of column 1."""
    o = nat.as_points(obs, "obs")
    if isinstance(texts, np.ndarray) and texts.ndim == 2:
        if texts.dtype != np.uint8 or not texts.flags.c_contiguous or nbytes is None:
            raise ValueError("a 2-D `texts` must be a C-contiguous uint8 array with `nbytes`")
        keep = texts
        B = len(nbytes)
        if B > texts.shape[0]:
            raise ValueError("more lengths than rows of `texts`")
        lens = np.ascontiguousarray(nbytes, dtype=np.int64)
        if B and (lens.min() < 0 or lens.max() > texts.shape[1]):
            raise ValueError("nbytes out of range")
        addr = np.uint64(texts.ctypes.data) + np.arange(B, dtype=np.uint64) * np.uint64(texts.strides[0])
    else:
        keep = [np.frombuffer(t, dtype=np.uint8) if not isinstance(t, np.ndarray) else np.ascontiguousarray(t).view(np.uint8).reshape(-1)
                for t in texts]
        B = len(keep)
        lens = np.array([k.shape[0] for k in keep], dtype=np.int64)
        addr = np.array([k.ctypes.data for k in keep], dtype=np.uint64)
    if isinstance(bandwidth, str):
        raise ValueError("a bandwidth rule needs the table sizes before the call; pass a number")
    common = _common(resolution, float(bandwidth), gamma, eps, log, kernel, algo)
    js, rows, status = np.empty(B), np.zeros(B, dtype=np.int64), np.zeros(B, dtype=np.int32)
    bnd = np.empty((B, 4)) if bounds else None
    if B:
        dev = default_device() if device is None else int(device)
        nat.check(nat.lib().kdejs_discrepancy_batch_text(
            dev, nat.ptr(o), o.shape[0], addr.ctypes.data_as(nat._c_dpp), lens.ctypes.data_as(nat._c_i64p), B,
            int(skip_rows), *common, nat.ptr(js), nat.ptr(rows), nat.ptr(status), nat.ptr(bnd)))
    del keep
    return (js, rows, status, bnd) if bounds else (js, rows, status)


def KDE_JS_divergence_pairs(sets, pairs, gamma=1.0, eps=1e-12, log=False, *,
                            resolution=_DEFAULT_RESOLUTION, bandwidth=_DEFAULT_BANDWIDTH,
                            kernel="epanechnikov", algo=None, device=None):
    """out[k] = KDE_JS_divergence(sets[i_k], sets[j_k]) for (i_k, j_k) in pairs; every set is
    uploaded once (the all-pairs loop of scripts/pairwise_2D_distance.py:112-118)."""
    arrays = [nat.as_points(s, f"sets[{i}]") for i, s in enumerate(sets)]
    pr = np.asarray(list(pairs), dtype=np.int32).reshape(-1, 2)
    ia = np.ascontiguousarray(pr[:, 0])
    ib = np.ascontiguousarray(pr[:, 1])
    common = _common(resolution, _bandwidth(bandwidth, *arrays), gamma, eps, log, kernel, algo)
    out = np.empty(len(pr), dtype=np.float64)
    ptrs, counts = nat.pointer_array(arrays)
    dev = default_device() if device is None else int(device)
    nat.check(nat.lib().kdejs_discrepancy_pairs(dev, ptrs, counts, len(arrays), nat.ptr(ia), nat.ptr(ib),
                                                len(pr), *common, nat.ptr(out)))
    return out


def summary_batch(obs, sims, sims_freqs, observed, gamma, core_threshold, rare_threshold, *, eps=0.0, log=False,
                  resolution=_DEFAULT_RESOLUTION, bandwidth=_DEFAULT_BANDWIDTH, algo=None, device=None):
    """The simulator's summary vectors for B simulations against one target, and the two ELFI nodes on top of them,
    in one device pass (kdejs_summary_batch): returns (summary (B,4) = [JS, freq_core, freq_intermediate, freq_rare],
    d (B,), log_d (B,)) -- process_result, poppunk_mod.py:351-382 with calculate_gene_freqs :285-291, then
    Distance("euclidean") against `observed` = [0, fc, fi, fr] (:494, :571) and Operation(np.log) (:572)."""
    o = nat.as_points(obs, "obs")
    sets = nat.PointSets(sims, "sims")
    freqs = [np.ascontiguousarray(np.asarray(f, dtype=np.float64).reshape(-1)) for f in sims_freqs]
    if len(freqs) != len(sets):
        raise ValueError("one gene-frequency vector per simulation is required")
    obs4 = np.ascontiguousarray(np.asarray(observed, dtype=np.float64).reshape(-1))
    if obs4.shape != (4,):
        raise ValueError("observed must be the 4-vector [0, freq_core, freq_intermediate, freq_rare]")
    common = _common(resolution, _bandwidth(bandwidth, o, sets), gamma, eps, log, "epanechnikov", algo)
    B = len(sets)
    summary, d, logd = np.empty((B, 4)), np.empty(B), np.empty(B)
    if B == 0:
        return summary, d, logd
    ptrs, counts, _, _keep = sets.pointers()
    fptrs = (ctypes.c_void_p * B)(*[f.ctypes.data for f in freqs])
    fcnt = (ctypes.c_int64 * B)(*[f.shape[0] for f in freqs])
    dev = default_device() if device is None else int(device)
    nat.check(nat.lib().kdejs_summary_batch(dev, nat.ptr(o), o.shape[0], ptrs, counts, fptrs, fcnt, B,
                                            float(core_threshold), float(rare_threshold), nat.ptr(obs4), *common,
                                            nat.ptr(summary), nat.ptr(d), nat.ptr(logd)))
    return summary, d, logd


def summary_from_js(js, sims_freqs, observed, core_threshold, rare_threshold, *, device=None):
    """summary_batch's last step for discrepancies that already exist (KDE_JS_divergence_batch_text: tables parsed on
    the device): the same kernel turns js (B,) and the B gene-frequency vectors into (summary (B,4), d, log_d)
    (kdejs_summary_from_js; poppunk_mod.py:285-291, :379, :571-572)."""
    js = np.ascontiguousarray(np.asarray(js, dtype=np.float64).reshape(-1))
    freqs = [np.ascontiguousarray(np.asarray(f, dtype=np.float64).reshape(-1)) for f in sims_freqs]
    B = js.shape[0]
    if len(freqs) != B:
        raise ValueError("one gene-frequency vector per simulation is required")
    obs4 = np.ascontiguousarray(np.asarray(observed, dtype=np.float64).reshape(-1))
    if obs4.shape != (4,):
        raise ValueError("observed must be the 4-vector [0, freq_core, freq_intermediate, freq_rare]")
    summary, d, logd = np.empty((B, 4)), np.empty(B), np.empty(B)
    if B == 0:
        return summary, d, logd
    fptrs = (ctypes.c_void_p * B)(*[f.ctypes.data for f in freqs])
    fcnt = (ctypes.c_int64 * B)(*[f.shape[0] for f in freqs])
    dev = default_device() if device is None else int(device)
    nat.check(nat.lib().kdejs_summary_from_js(dev, nat.ptr(js), fptrs, fcnt, B, float(core_threshold),
                                              float(rare_threshold), nat.ptr(obs4), nat.ptr(summary), nat.ptr(d),
                                              nat.ptr(logd)))
    return summary, d, logd


def kde_density(df_scaled, resolution=_DEFAULT_RESOLUTION, bandwidth=_DEFAULT_BANDWIDTH,
                kernel="epanechnikov", minimum=0.0, maximum=1.0, *, algo=None, device=None):
    """exp(KernelDensity(bandwidth, kernel).fit(X).score_samples(get_grid(min, max, R)[2])) as an
    (R, R) array (row = column-0 node): get_kde + generate_samples, KDE_distance.py:48-64, for
    points that are already scaled; also the density plot_scatter draws (poppunk_mod.py:98-105)."""
    X = nat.as_points(df_scaled, "df_scaled")
    if not np.isfinite(X).all():
        raise ValueError("Input X contains NaN or infinity.")
    R, h, _, _, _, k, al = _common(resolution, _bandwidth(bandwidth, X), 1.0, 0.0, False, kernel, algo)
    out = np.empty(R * R, dtype=np.float64)
    dev = default_device() if device is None else int(device)
    nat.check(nat.lib().kdejs_kde_density(dev, nat.ptr(X), X.shape[0], float(minimum), float(maximum), R, h, k,
                                          al, nat.ptr(out)))
    return out.reshape(R, R)


def _plot_contours(df1, df2, z1, z2, eps, log, resolution, outplot):
    """The two contour PNGs of KDE_distance.py:110-126 (host-side rendering only)."""
    import matplotlib.pyplot as plt    # optional dependency, exactly as in the reference

    a, b = np.asarray(df1, dtype=np.float64), np.asarray(df2, dtype=np.float64)
    if log:
        a, b = np.log(a + eps), np.log(b + eps)
    both = np.vstack([a, b])
    lo, hi = np.nanmin(both, axis=0), np.nanmax(both, axis=0)
    span = np.where(hi - lo < 10 * np.finfo(float).eps, 1.0, hi - lo)
    xx, yy, _ = get_grid(0, 1, resolution)
    for pts, z, tag in ((a, z1, "_z1_contours.png"), (b, z2, "_z2_contours.png")):
        sc = np.nan_to_num((pts - lo) / span, nan=0.0)
        sq = z.reshape(xx.shape).T
        plt.figure(figsize=(11, 8), dpi=160, facecolor="w", edgecolor="k")
        levels = np.linspace(sq.min(), sq.max(), 100)
        plt.contour(xx, yy, sq, levels=levels[1:], cmap="plasma")
        plt.scatter(sc[:, 0], sc[:, 1], s=1, alpha=1)
        plt.savefig(outplot + tag)
        plt.close()


def get_options(argv=None):
    parser = argparse.ArgumentParser(description="Fit KDE to two 2D distributions and generates a distance.",
                                     prog="python KDE_distance.py")
    io = parser.add_argument_group("Input/Output options")
    io.add_argument("--infile1", required=True, help="Input distance file 1.")
    io.add_argument("--infile2", required=True, help="Input distance file 2.")
    io.add_argument("--gamma", default=0.25, type=float,
                    help="Power to raise KDE distributions to for smoothing (Default = 0.25).")
    io.add_argument("--plot-outpref", default=None, type=str,
                    help="Prefix for output plots. If not provided, no plots generated.")
    io.add_argument("--resolution", default=_DEFAULT_RESOLUTION, type=int, help="Grid nodes per axis (Default = 100).")
    return parser.parse_args(argv)


def main(argv=None):
    opt = get_options(argv)
    # same values np.loadtxt(..., delimiter='\t') gives (KDE_distance.py:138-139), parsed by the library's
    # multi-threaded reader straight into pinned memory
    df1 = nat.TsvTable(opt.infile1, skip_rows=0).array
    df2 = nat.TsvTable(opt.infile2, skip_rows=0).array
    js = KDE_JS_divergence(df1, df2, gamma=opt.gamma, eps=0.0, log=False, outplot=opt.plot_outpref,
                           resolution=opt.resolution)
    print(f"js_distance_no_eps: {js}")
    if opt.gamma < 1.0:
        # stdout keeps the reference's single line (KDE_distance.py:144); the caveat goes to stderr
        print(f"note: exact sum; upstream's value at gamma={opt.gamma:g} carries its ball tree's residue at "
              "zero-density nodes (~5e-4 relative at 0.25, up to ~10 % at 1e-12; INTEGRATION.md section 5)",
              file=sys.stderr)


if __name__ == "__main__":
    main()
