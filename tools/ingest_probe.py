"""Tables on disk -> discrepancies: the sweep the reference runs with np.loadtxt + one CPU evaluation per file
(scripts/distance_to_gridsearch.py:48-55, :109-125).  Parse-only rates and sweeps.gridsearch_best end to end.
usage: python tools/ingest_probe.py [n_files]      (GPU box; writes its tables under a temporary directory)"""
import os, shutil, sys, tempfile, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.getcwd())
import numpy as np
import poppunk_mod_b200 as pk
from poppunk_mod_b200 import sweeps
from oracle import workloads as wl

nat = pk._native
n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = tempfile.mkdtemp(prefix="kdejs_ingest_")
try:
    T = wl.example_target()
    ref = os.path.join(d, "ref.tsv")
    with open(ref, "w") as f:
        f.writelines(f"{x!r}\t{y!r}\n" for x, y in T.tolist())          # the reference's own spelling: repr()
    distinct = []
    for s in range(32):
        q = os.path.join(d, f"sim{s}.tsv")
        with open(q, "w") as f:
            a = wl.sim_batch_member(s)
            a[:, 0] = np.round(a[:, 0], 6)          # core distances as the simulator prints them (the example target: 0.000917)
            f.writelines(f"{x!r}\t{y!r}\n" for x, y in a.tolist())
        distinct.append(q)
    files = []
    for k in range(n_files):
        q = os.path.join(d, f"run{k}.tsv")
        if k < 32:
            os.rename(distinct[k], q)
            distinct[k] = q
        else:
            shutil.copy(distinct[k % 32], q)
        files.append(q)
    mb = sum(os.path.getsize(q) for q in files) / 1e6
    print(f"{n_files} tables of 100 000 rows, {mb / n_files:.2f} MB each; {os.cpu_count()} host cores", flush=True)
    t0 = time.perf_counter(); a = np.loadtxt(files[0], delimiter="\t"); t_np = time.perf_counter() - t0
    print(f"np.loadtxt (the reference's reader): {1e3 * t_np:.1f} ms per table = {1 / t_np:.0f} tables/s per core", flush=True)
    for th in (1, 2, 4, 8):
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); t = nat.TsvTable(files[1], pinned=False, threads=th); ts.append(time.perf_counter() - t0)
        print(f"kdejs_tsv_load, {th} thread(s): {1e3 * min(ts):.2f} ms per table ({mb / n_files / min(ts) / 1e3:.2f} GB/s)", flush=True)
    assert np.array_equal(nat.TsvTable(files[0], pinned=False).array, a)
    arena = nat.TableArena(64, 100000)
    arena.load(files[:64], workers=8)                                     # first touch of the arena
    for w in (4, 8, 16, 32):
        t0 = time.perf_counter()
        for c in range(0, n_files, 64):
            arena.load(files[c:c + 64], skip_rows=0, workers=w)
        dt = time.perf_counter() - t0
        print(f"parse into a page-locked arena, {w} files at a time: {n_files / dt:.0f} tables/s ({mb / dt / 1e3:.2f} GB/s)", flush=True)
    arena.close()
    for pinned in (True, False):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=16) as ex:
            for c in range(0, n_files, 64):
                tabs = list(ex.map(lambda p: nat.TsvTable(p, skip_rows=0, pinned=pinned, threads=1), files[c:c + 64]))
                for t in tabs:
                    t.close()
        dt = time.perf_counter() - t0
        print(f"one {'page-locked' if pinned else 'ordinary'} allocation per table, 16 files at a time: {n_files / dt:.0f} tables/s", flush=True)
    text = nat.TextArena(64, int(os.path.getsize(files[0]) * 1.5))
    for w in (2, 4, 8, 16):
        t0 = time.perf_counter()
        for c in range(0, n_files, 64):
            text.read(files[c:c + 64], threads=w)
        dt = time.perf_counter() - t0
        print(f"bytes into a page-locked arena (no parsing), {w} files at a time: {n_files / dt:.0f} tables/s ({mb / dt / 1e3:.2f} GB/s)", flush=True)
    nb = text.read(files[:64], threads=8).copy()
    for _ in range(2):
        t0 = time.perf_counter()
        js, rows, status = pk.KDE_JS_divergence_batch_text(T, text.array, 0.25, 0.0, nbytes=nb, skip_rows=0)
        dt = time.perf_counter() - t0
    print(f"kdejs_discrepancy_batch_text, 64 tables already in page-locked memory (upload + device parse + evaluate, "
          f"100x100 grid): {64 / dt:.0f} tables/s ({dt * 1e3:.1f} ms)", flush=True)
    assert (status == 0).all()
    import ctypes
    L = nat.lib()
    nat.check(L.kdejs_profile_reset(0)); nat.check(L.kdejs_profile_enable(0, 1))
    for _ in range(3):
        pk.KDE_JS_divergence_batch_text(T, text.array, 0.25, 0.0, nbytes=nb, skip_rows=0, bounds=True)
    pm, pn = ctypes.c_double(), ctypes.c_int64()
    tb = 3 * float(nb.sum())
    for which, name, traffic in ((10, "k_tsv_count + k_tsv_scan", tb), (11, "k_tsv_parse + k_tsv_bounds", tb + 3 * 2 * 16.0 * float(rows.sum()))):
        nat.check(L.kdejs_profile_read(0, which, ctypes.byref(pm), ctypes.byref(pn)))
        print(f"   {name}: {1e3 * pm.value / 3 / 64:.1f} us per table, {traffic / (pm.value * 1e-3) / 1e9:.0f} GB/s of algorithmic "
              f"traffic (text once per kernel; rows written and read once)", flush=True)
    nat.check(L.kdejs_profile_enable(0, 0))
    text.close()
    ref_dist = None
    for parse, io, chunk in (("host", 16, 64), ("device", 4, 64), ("device", 8, 64), ("device", 16, 64), ("device", 8, 32),
                            ("device", 16, 128), ("device", 16, 256)):
        kw = dict(gamma=0.25, eps=0.0, log=False, io_threads=io, chunk=chunk, parse=parse)
        sweeps.gridsearch_best(ref, files[:2 * chunk], **kw)      # warm-up (arenas are kept between calls)
        t0 = time.perf_counter()
        best, dist = sweeps.gridsearch_best(ref, files, **kw)
        dt = time.perf_counter() - t0
        print(f"sweeps.gridsearch_best, files -> arg-min, parse on the {parse}, io_threads={io}, {chunk} files per batch: "
              f"{n_files / dt:.0f} tables/s ({dt:.2f} s; best {os.path.basename(best[0])} d={best[1]:.6f})", flush=True)
        if ref_dist is None:
            ref_dist = dist
        assert np.array_equal(dist, ref_dist)
    one = pk.KDE_JS_divergence(T, a, 0.25, 0.0)
    assert dist[0] == one, (dist[0], one)
finally:
    shutil.rmtree(d, ignore_errors=True)
