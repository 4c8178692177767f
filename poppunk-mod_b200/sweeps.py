"""Batch drivers: the two sweep scripts of the reference on top of the batched GPU entry points.

    pairwise_distances   scripts/pairwise_2D_distance.py:47-79, :100-118  (all pairs of species files,
                         optional euclidean combination with gene-class fractions)
    gridsearch_best      scripts/distance_to_gridsearch.py:48-55, :75-125 (one reference against every
                         simulation of a grid search; running arg-min under optional bounds)
    read_file / read_distfile   the scripts' and poppunk_mod.py's table readers (:38-45, :255-269)

Files are parsed by the library's multi-threaded reader straight into pinned memory while the GPU works
on the previous chunk; the discrepancies come from KDE_JS_divergence_pairs / _batch.  Command line:

    python -m poppunk_mod_b200.sweeps pairwise --infile table.txt --outpref out
    python -m poppunk_mod_b200.sweeps gridsearch --indir runs/ --ref obs.txt
"""
from __future__ import annotations

import argparse
import itertools
import math
import os
import sys
import threading
import warnings
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

from . import _native as nat
from . import KDE_distance as kd


def read_file(path, pinned=None, threads=0):
    """Last two columns of a headerless tab-separated table (pairwise_2D_distance.py:38-45)."""
    return nat.TsvTable(path, skip_rows=0, pinned=pinned, threads=threads)


def read_distfile(path, pinned=None, threads=0):
    """poppunk_mod.read_distfile (:255-269): a 2-column file is read whole; otherwise pandas' inferred
    header consumes the first line and the first two (id) columns are dropped."""
    with open(path, "r") as f:
        two_column = len(f.readline().rstrip().split("\t")) == 2
    return nat.TsvTable(path, skip_rows=0 if two_column else 1, pinned=pinned, threads=threads)


def _bounds(a):
    return (np.min(a[:, 0]), np.max(a[:, 0]), np.min(a[:, 1]), np.max(a[:, 1]))


def _load_many(paths, reader, workers):
    """Parse many tables concurrently, `workers` files at a time with one parser thread each, into ordinary memory
    (the library stages pageable uploads itself; a page-locked allocation per table costs more than parsing it:
    measured on a 16-core host, 100 000-row tables: 150-340 tables/s with one allocation each, 1 350-1 740 into
    slots of one page-locked arena -- gridsearch_best -- against 30 per core for np.loadtxt)."""
    with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:      # ctypes releases the GIL while parsing
        return list(ex.map(lambda p: reader(p, pinned=False, threads=1), paths))


def pairwise_distances(files, gamma=0.25, eps=0.0, log=False, *, io_threads=8, device=None, resolution=100):
    """files: list of paths, or of (path, core_frac, intermediate_frac, rare_frac) tuples.
    Returns [(file1, file2, distance, (min_core, max_core, min_acc, max_acc) of file2), ...] in
    itertools.combinations order.  With fractions the distance is the euclidean combination the
    reference builds (pairwise_2D_distance.py:60-77): dist((0,c1,i1,r1), (JS,c2,i2,r2))."""
    entries = [(f,) if isinstance(f, (str, Path)) else tuple(f) for f in files]
    tables = _load_many([str(e[0]) for e in entries], read_file, io_threads)
    sets = [t.array for t in tables]
    pairs = list(itertools.combinations(range(len(entries)), 2))
    js = kd.KDE_JS_divergence_pairs(sets, pairs, gamma, eps, log, device=device, resolution=resolution)
    bounds = [_bounds(s) for s in sets]
    out = []
    for (i, j), d in zip(pairs, js):
        e1, e2 = entries[i], entries[j]
        dist = d if len(e1) == 1 else math.dist((0, e1[1], e1[2], e1[3]), (d, e2[1], e2[2], e2[3]))
        out.append((str(e1[0]), str(e2[0]), dist, bounds[j]))
    for t in tables:
        t.close()
    return out


def _in(value, bounds):
    return bounds[0] <= value <= bounds[1]


_SMALL_GAMMA_NOTE = (
    "kdejs returns the exact sum; with gamma << 1 upstream's value is dominated by its ball tree's residue at "
    "zero-density nodes, so distances (up to ~10 %) and the file kept as best can differ from upstream's "
    "(INTEGRATION.md section 5: arg-min differed, 5 % of pairs reordered on 200 simulations at gamma=1e-12)")


# Page-locked arenas are kept between sweeps (allocating 100+ MB of page-locked memory takes ~0.1 s, as long as a
# short sweep itself): key -> [capacity, arenas, in use].  A sweep that finds its key in use (another thread sweeping
# on the same device) works with arenas of its own and frees them afterwards.  release_arenas() gives the memory back.
_ARENAS = {}
_ARENAS_LOCK = threading.Lock()


class _ArenaLease:
    def __init__(self, kind, count, slots, capacity, make):
        self.key, self.count, self.slots, self.capacity, self.make = (kind, slots), count, slots, capacity, make
        self.private = None

    def __enter__(self):
        with _ARENAS_LOCK:
            have = _ARENAS.get(self.key)
            if have is not None and have[2]:
                self.private = []                    # the cached arenas are busy: this sweep gets its own
            else:
                if have is None or have[0] < self.capacity:
                    if have is not None:
                        for a in have[1]:
                            a.close()
                    _ARENAS[self.key] = have = [self.capacity, [], False]
                have[2] = True
        if self.private is not None:
            self.private = [self.make(self.slots, self.capacity) for _ in range(self.count)]
            return self.private
        while len(have[1]) < self.count:             # a longer sweep than the last one: only the missing arena is added
            have[1].append(self.make(self.slots, have[0]))
        return have[1][:self.count]

    def __exit__(self, *exc):
        if self.private is not None:
            for a in self.private:
                a.close()
        else:
            with _ARENAS_LOCK:
                _ARENAS[self.key][2] = False
        return False


def release_arenas():
    """Frees the page-locked arenas the sweep drivers keep between calls (up to ~0.5 GB per device)."""
    with _ARENAS_LOCK:
        for key in [k for k, v in _ARENAS.items() if not v[2]]:
            for a in _ARENAS.pop(key)[1]:
                a.close()


def _accept(best, name, d, b, core_min_bounds, core_max_bounds, acc_min_bounds, acc_max_bounds):
    """distance_to_gridsearch.py:112-125: a smaller distance replaces the best one if its table's bounds pass."""
    if not d < best[1]:
        return best
    ok = True
    if core_min_bounds is not None and core_max_bounds is not None:
        ok &= _in(b[0], core_min_bounds) and _in(b[1], core_max_bounds)
    if acc_min_bounds is not None and acc_max_bounds is not None:
        ok &= _in(b[2], acc_min_bounds) and _in(b[3], acc_max_bounds)
    return (name, d, tuple(b)) if ok else best


def _device_parse_worker(ref_arr, chunks, gamma, eps, log, chunk, io_threads, device, resolution):
    """One device's share of a sweep with NO parser on the host: the bytes of the next chunk of files are read into a
    page-locked arena (kdejs_files_read) while the GPU uploads, parses and evaluates the current one
    (kdejs_discrepancy_batch_text).  Tables the device parser does not cover (missing-value spellings, inf / nan, more
    than 19 digits, files larger than a slot) go through the host reader.  Returns [(distances, bounds)] per chunk."""
    out = []
    if not chunks:
        return out
    slot_bytes = int(os.path.getsize(chunks[0][0]) * 1.5) + 4096
    with _ArenaLease(("text", device), min(2, len(chunks)), chunk, slot_bytes, lambda n, cap: nat.TextArena(n, cap)) as arenas, \
            ThreadPoolExecutor(max_workers=1) as prefetch:
        read = lambda ci: arenas[ci % 2].read(chunks[ci], threads=io_threads).copy()
        pending = prefetch.submit(read, 0)
        for ci, names in enumerate(chunks):
            nbytes = pending.result()
            pending = prefetch.submit(read, ci + 1) if ci + 1 < len(chunks) else None
            js, rows, status, bnd = kd.KDE_JS_divergence_batch_text(
                ref_arr, arenas[ci % 2].array, gamma, eps, log, nbytes=np.maximum(nbytes, 0), skip_rows=0,
                resolution=resolution, device=device, bounds=True)
            redo = [k for k in range(len(names)) if nbytes[k] < 0 or status[k] == nat.KDEJS_TABLE_HOST]
            if redo:        # the host reader decides (and reports malformed rows the way it always did)
                tables = _load_many([names[k] for k in redo], read_file, io_threads)
                js[redo] = kd.KDE_JS_divergence_batch(ref_arr, [t.array for t in tables], gamma, eps, log,
                                                      devices=device, resolution=resolution)
                for k, t in zip(redo, tables):
                    bnd[k] = _bounds(t.array)
            inf = [names[k] for k in range(len(names)) if status[k] == nat.KDEJS_TABLE_INF and k not in redo]
            if inf:
                raise ValueError(f"Input X contains infinity or a value too large for dtype('float64'). ({inf[0]})")
            out.append((js, bnd))
    return out


def _gridsearch_device_parse(ref_arr, files, gamma, eps, log, bounds_kw, chunk, io_threads, devices, resolution):
    """Chunks of `chunk` files are dealt to the devices round-robin, one host thread per device (the ctypes calls
    release the GIL); the running arg-min of distance_to_gridsearch.py:112-125 is then taken in file order, so the
    result does not depend on the number of devices."""
    chunks = [files[i:i + chunk] for i in range(0, len(files), chunk)]
    n = max(1, min(len(devices), len(chunks)))
    per_dev_threads = max(2, io_threads // n)
    shares = [chunks[k::n] for k in range(n)]
    if n == 1:
        results = [_device_parse_worker(ref_arr, shares[0], gamma, eps, log, chunk, io_threads, devices[0], resolution)]
    else:
        with ThreadPoolExecutor(max_workers=n) as ex:
            futs = [ex.submit(_device_parse_worker, ref_arr, shares[k], gamma, eps, log, chunk, per_dev_threads,
                              devices[k], resolution) for k in range(n)]
            results = [f.result() for f in futs]
    best = (None, 1.0)
    distances = np.empty(len(files))
    pos = 0
    for ci, names in enumerate(chunks):
        js, bnd = results[ci % n][ci // n]
        for name, d, b in zip(names, js, bnd):
            distances[pos] = d
            pos += 1
            best = _accept(best, name, d, b, **bounds_kw)
    return best, distances


def gridsearch_best(ref, files, gamma=1e-12, eps=1e-12, log=True, *, core_min_bounds=None, core_max_bounds=None,
                    acc_min_bounds=None, acc_max_bounds=None, chunk=64, io_threads=None, devices=None,
                    resolution=100, parse=None):
    """One reference against many simulations.  Defaults are the reference's call
    KDE_JS_divergence(df1, df2, 1e-12, log=True) (distance_to_gridsearch.py:50).  Bounds are absolute
    [lo, hi] pairs (the script multiplies its fractions by the reference's column maxima, :89-99).
    Returns (best, distances) with best = (file, distance, bounds) or (None, 1.0), distances in file order.

    parse: "device" (default) -- the files' bytes go to the GPU(s) and are parsed there, chunks dealt to `devices`
    round-robin; "host" (or KDEJS_PARSE=host) -- the library's host reader parses them into page-locked arenas."""
    ref_arr = ref if isinstance(ref, np.ndarray) else read_file(ref).array
    files = [str(f) for f in files]
    if io_threads is None:          # one reader / parser thread per file; measured on a 16-core host: 560 tables/s with 8, 1170 with 16
        io_threads = max(1, min(16, os.cpu_count() or 1))
    if gamma < 0.05:
        warnings.warn(_SMALL_GAMMA_NOTE, stacklevel=2)
    best = (None, 1.0)
    distances = np.empty(len(files))
    if not files:
        return best, distances
    if parse is None:
        parse = os.environ.get("KDEJS_PARSE", "device")
    if parse not in ("device", "host"):
        raise ValueError("parse must be 'device' or 'host'")
    bounds_kw = dict(core_min_bounds=core_min_bounds, core_max_bounds=core_max_bounds, acc_min_bounds=acc_min_bounds,
                     acc_max_bounds=acc_max_bounds)
    if parse == "device":
        devs = [kd.default_device()] if devices is None else [devices] if isinstance(devices, int) else list(devices)
        return _gridsearch_device_parse(ref_arr, files, gamma, eps, log, bounds_kw, chunk, io_threads, devs, resolution)
    chunks = [files[i:i + chunk] for i in range(0, len(files), chunk)]
    # Two page-locked arenas of `chunk` slots: the files of chunk k+1 are parsed straight into their slots (one
    # parser thread per file, io_threads files at a time) while the GPU works on chunk k; nothing is allocated per
    # file.  Slots hold 1.5 x the rows of the first file; a longer table is read into memory of its own.
    first = read_file(files[0], pinned=False)
    cap = int(first.rows * 1.5) + 1024
    first.close()
    with _ArenaLease("rows", min(2, len(chunks)), chunk, cap, lambda n, c: nat.TableArena(n, c)) as arenas, \
            ThreadPoolExecutor(max_workers=1) as prefetch, ThreadPoolExecutor(max_workers=max(1, io_threads)) as io:
        load = lambda ci: arenas[ci % 2].load(chunks[ci], skip_rows=0, executor=io, overflow="alloc")   # np.loadtxt, :49
        pending = prefetch.submit(load, 0)
        pos = 0
        for ci, names in enumerate(chunks):
            tables = pending.result()
            # chunk ci+1 is parsed into the OTHER arena (chunk ci-1's, whose batch has returned) during this batch
            pending = prefetch.submit(load, ci + 1) if ci + 1 < len(chunks) else None
            js = kd.KDE_JS_divergence_batch(ref_arr, tables, gamma, eps, log, devices=devices, resolution=resolution)
            for name, t, d in zip(names, tables, js):
                distances[pos] = d
                pos += 1
                if d < best[1]:                                     # :112-125
                    best = _accept(best, name, d, _bounds(t), **bounds_kw)
    return best, distances


def _pairwise_main(o):
    if o.indir is None and o.infile is None:
        print("Please specify one of --indir or --infile")
        return 1
    if o.indir is not None:
        files = [str(p) for p in Path(o.indir).glob("*.txt")]
    else:
        files = []
        with open(o.infile) as f:
            f.readline()
            for line in f:
                s = line.rstrip().split("\t")
                pan = int(s[1])
                files.append((s[6], int(s[2]) / pan, int(s[3]) / pan, int(s[4]) / pan))
    rows = pairwise_distances(files, io_threads=o.threads)
    with open(o.outpref + ".tsv", "w") as out:
        out.write("File1\tFile2\tDistance\n")
        for f1, f2, d, _ in rows:
            out.write(f"{f1}\t{f2}\t{d}\n")
    return 0


def _gridsearch_main(o):
    files = [str(p) for p in Path(o.indir).glob("**/*.tsv")]
    ref = read_file(o.ref).array
    mc, ma = np.max(ref[:, 0]), np.max(ref[:, 1])
    print(f"max_core_ref: {mc}, max_acc_ref: {ma}")
    frac = lambda s, m: None if s is None else [float(x) * m for x in s.split(",")]
    kw = dict(core_min_bounds=frac(o.core_min_bounds, mc), core_max_bounds=frac(o.core_max_bounds, mc),
              acc_min_bounds=frac(o.acc_min_bounds, ma), acc_max_bounds=frac(o.acc_max_bounds, ma))
    best, _ = gridsearch_best(ref, files, io_threads=o.threads, **kw)
    if len(best) == 2:
        print("No best fitting distribution found.")
    else:
        b = best[2]
        print(f"File: {best[0]}, Min distance: {best[1]}, Min core: {b[0]}, Max core: {b[1]}, Min acc: {b[2]}, Max acc: {b[3]}")
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m poppunk_mod_b200.sweeps")
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("pairwise", help="scripts/pairwise_2D_distance.py on the GPU")
    p.add_argument("--indir", default=None)
    p.add_argument("--infile", default=None)
    p.add_argument("--outpref", default="output")
    p.add_argument("--threads", type=int, default=8)
    g = sub.add_parser("gridsearch", help="scripts/distance_to_gridsearch.py on the GPU")
    g.add_argument("--indir", required=True)
    g.add_argument("--ref", required=True)
    g.add_argument("--outpref", default="output")
    for name in ("core-min-bounds", "core-max-bounds", "acc-min-bounds", "acc-max-bounds"):
        g.add_argument("--" + name, type=str, default=None)
    g.add_argument("--threads", type=int, default=None)
    o = ap.parse_args(argv)
    return _pairwise_main(o) if o.cmd == "pairwise" else _gridsearch_main(o)


if __name__ == "__main__":
    sys.exit(main())
