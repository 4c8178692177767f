"""All pairs of species tables (scripts/pairwise_2D_distance.py:100-118: itertools.combinations over the files, one
KDE_JS_divergence(gamma=0.25, eps=0) per pair on a 100x100 grid, optional euclidean combination :60-77):
sweeps.pairwise_distances from files, and KDE_JS_divergence_pairs on the parsed arrays alone.
usage: python tools/pairwise_probe.py [n_species]      (GPU box; writes its tables under a temporary directory)"""
import itertools, os, shutil, sys, tempfile, time
sys.path.insert(0, os.getcwd())
import numpy as np
import poppunk_mod_b200 as pk
from poppunk_mod_b200 import sweeps
from oracle import workloads as wl

S = int(sys.argv[1]) if len(sys.argv) > 1 else 96
d = tempfile.mkdtemp(prefix="kdejs_pairwise_")
try:
    rng = np.random.default_rng(7)
    sizes = rng.integers(20706, 100001, S)               # the row counts of the reference's distances/2_column tables
    files, sets = [], []
    for s in range(S):
        a = wl.sim_batch_member(s, n=int(sizes[s]))
        a[:, 0] = np.round(a[:, 0], 6)
        q = os.path.join(d, f"species{s}.txt")
        with open(q, "w") as f:
            f.writelines(f"{x!r}\t{y!r}\n" for x, y in a.tolist())
        files.append(q)
        sets.append(a)
    pairs = list(itertools.combinations(range(S), 2))
    print(f"{S} species tables of {sizes.min()}-{sizes.max()} rows ({sizes.sum()} in all), {len(pairs)} pairs; "
          f"{os.cpu_count()} host cores", flush=True)
    sweeps.pairwise_distances(files[:8])                                        # warm-up
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        rows = sweeps.pairwise_distances(files, io_threads=16)
        best = min(best, time.perf_counter() - t0)
    print(f"sweeps.pairwise_distances, files -> table of distances: {len(pairs) / best:.0f} pairs/s ({best:.3f} s)", flush=True)
    best_a = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        js = pk.KDE_JS_divergence_pairs(sets, pairs, 0.25, 0.0)
        best_a = min(best_a, time.perf_counter() - t0)
    print(f"KDE_JS_divergence_pairs on the parsed arrays (upload of every set once + evaluation): "
          f"{len(pairs) / best_a:.0f} pairs/s ({best_a:.3f} s)", flush=True)
    assert np.array_equal(js, np.array([r[2] for r in rows]))
    for k in (0, len(pairs) // 3, len(pairs) - 1):
        i, j = pairs[k]
        assert js[k] == pk.KDE_JS_divergence(sets[i], sets[j], 0.25, 0.0), k
    t0 = time.perf_counter()
    one = pk.KDE_JS_divergence(sets[0], sets[1], 0.25, 0.0, algo="binned")
    print(f"driver = pairs call = single calls bit for bit; the reference: one pair per core in ~0.3 s "
          f"(np.loadtxt twice + sklearn at R = 100)", flush=True)
finally:
    shutil.rmtree(d, ignore_errors=True)
