"""Simulator post-processing around the discrepancy (the caller side of the hot path).

Mirrors, for arrays already in memory or files on disk:

    process_result         poppunk_mod.py:351-382   -> [divergence, freq_core, freq_intermediate, freq_rare]
    the ELFI nodes built on it: Distance("euclidean", sim) and Operation(np.log, d)   poppunk_mod.py:571-572

The whole summary runs on the device behind the Jensen-Shannon reduction (KDE_distance.summary_batch ->
kdejs_summary_batch): the gene-frequency classes of calculate_gene_freqs (poppunk_mod.py:285-291), the euclidean
distance to the observed vector [0, fc, fi, fr] (:494) and its log come back with the discrepancies in one copy.
process_results_batch does a whole batch of simulations (BOLFI initial evidence, a Latin-hypercube grid) with ONE
call; the target is read once instead of once per simulation (poppunk_mod.py:371 re-reads it every time).
"""
from __future__ import annotations

import os

import numpy as np

from . import KDE_distance as kd
from . import _native as nat
from .sweeps import _device_parse_worker, read_distfile


def _points(x):
    if isinstance(x, (str, os.PathLike)):
        return read_distfile(x).array
    return x


def _freqs(x):
    if isinstance(x, (str, os.PathLike)):
        try:                    # the library's reader: the same values, several times faster, and it releases the GIL
            return nat.load_numbers(x)
        except ValueError:      # anything it does not read (e.g. another delimiter): what the reference does, :361
            return np.loadtxt(x, delimiter="\t", dtype="float64")
    return np.asarray(x, dtype=np.float64)


def _two_column(path):
    """The test of read_distfile (poppunk_mod.py:255-269): such a file is read whole, as np.loadtxt reads it (:360)."""
    with open(path, "rb") as f:
        return f.readline().rstrip().count(b"\t") == 1


def _all_two_column(paths):
    """One open + one short read per file, on the calling thread: 256 files take 2 ms; handing them to a thread pool
    was measured 8x slower (the work is system calls and interpreter time, and the pool's threads take turns on the GIL)."""
    return all(_two_column(p) for p in paths)


def _from_files_on_device(obs_arr, files, freq_files, gamma, core_threshold, rare_threshold, observed, device,
                          resolution, chunk=64):
    """Simulations on disk without a parser on the host: the tables' BYTES go to the GPU in chunks and are parsed and
    evaluated there (sweeps._device_parse_worker -> kdejs_discrepancy_batch_text; tables the device parser does not
    cover go through the host reader, exactly as in the grid-search sweep) while host threads read the small
    gene-frequency files; the summary kernel then runs on the discrepancies (kdejs_summary_from_js).  The same
    doubles reach the same kernels, so the vectors equal the host-parsed path's bit for bit."""
    from concurrent.futures import ThreadPoolExecutor

    dev = kd.default_device() if device is None else int(device)
    chunks = [files[i:i + chunk] for i in range(0, len(files), chunk)]
    io_threads = max(2, min(16, os.cpu_count() or 1))
    with ThreadPoolExecutor(max_workers=4) as ex:
        freq_futures = [ex.submit(_freqs, f) for f in freq_files]
        per_chunk = _device_parse_worker(obs_arr, chunks, gamma, 0.0, False, chunk, io_threads, dev, resolution)
        freqs = [f.result() for f in freq_futures]
    js = np.concatenate([r[0] for r in per_chunk])
    return kd.summary_from_js(js, freqs, observed, core_threshold, rare_threshold, device=dev)


def process_results_batch(simulations, simulations_freqs, obs, gamma, core_threshold, rare_threshold, *,
                          observed=None, device=None, resolution=100, algo=None, parse=None):
    """(B,4) summary vectors for B simulations against one target.  With `observed` (the 4-vector of
    poppunk_mod.py:494) also returns the ELFI nodes d and log_d: (summary, d, log_d).

    parse (simulations given as files): "device" (default; KDEJS_PARSE overrides) -- two-column tables are parsed on
    the GPU; "host" -- the library's host reader parses them, one thread per file."""
    obs_arr = _points(obs)
    simulations, simulations_freqs = list(simulations), list(simulations_freqs)
    if len(simulations) != len(simulations_freqs):
        raise ValueError("one gene-frequency vector per simulation is required")
    want_d = observed is not None
    obs4 = observed if want_d else [0.0, 0.0, 0.0, 0.0]
    if parse is None:
        parse = os.environ.get("KDEJS_PARSE", "device")
    if parse not in ("device", "host"):
        raise ValueError("parse must be 'device' or 'host'")
    on_disk = len(simulations) > 1 and all(isinstance(s, (str, os.PathLike)) for s in simulations)
    if on_disk and parse == "device" and algo is None and _all_two_column(simulations):
        summary, d, log_d = _from_files_on_device(obs_arr, [os.fspath(s) for s in simulations], simulations_freqs, gamma,
                                                  core_threshold, rare_threshold, obs4, device, resolution)
        return (summary, d, log_d) if want_d else summary
    if on_disk:
        # tables on disk: parsed concurrently, one parser thread per file (ctypes releases the GIL); ordinary memory,
        # which the library stages itself -- a page-locked allocation per file costs more than parsing it
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=max(1, min(16, os.cpu_count() or 1))) as ex:
            sims = list(ex.map(lambda p: read_distfile(p, pinned=False, threads=1).array, simulations))
            freqs = list(ex.map(_freqs, simulations_freqs))
    else:
        sims = [_points(s) for s in simulations]
        freqs = [_freqs(f) for f in simulations_freqs]
    summary, d, log_d = kd.summary_batch(obs_arr, sims, freqs, obs4, gamma,
                                         core_threshold, rare_threshold, eps=0, log=False, resolution=resolution,
                                         algo=algo, device=device)
    return (summary, d, log_d) if want_d else summary


def process_result(simulations, simulations_freqs, obs, gamma, core_threshold, rare_threshold, *, device=None,
                   resolution=100, algo=None):
    """One simulation: `simulations` / `obs` are (N,2) arrays or TSV paths, `simulations_freqs` a vector or path."""
    return process_results_batch([simulations], [simulations_freqs], obs, gamma, core_threshold, rare_threshold,
                                 device=device, resolution=resolution, algo=algo)[0]


def log_euclidean_discrepancy(summaries, observed):
    """log of the euclidean distance between summary vectors and the observed one ([0, fc, fi, fr],
    poppunk_mod.py:494): the quantity BOLFI's GP is fitted to, for summaries that are already on the host
    (process_results_batch(observed=...) returns it straight from the device)."""
    s = np.atleast_2d(np.asarray(summaries, dtype=np.float64))
    o = np.asarray(observed, dtype=np.float64)
    return np.log(np.sqrt(((s - o) ** 2).sum(axis=1)))
