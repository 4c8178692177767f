"""Simulations on disk -> summary vectors and ELFI nodes (process_result, poppunk_mod.py:351-382, for a batch):
simulator.process_results_batch with the tables parsed by the host reader vs on the device.
usage: python tools/simulator_probe.py [n_files]      (GPU box; writes its tables under a temporary directory)"""
import os, shutil, sys, tempfile, time
sys.path.insert(0, os.getcwd())
import numpy as np
import poppunk_mod_b200 as pk
from poppunk_mod_b200 import simulator
from oracle import workloads as wl

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = tempfile.mkdtemp(prefix="kdejs_simulator_")
try:
    T = wl.example_target()
    rng = np.random.default_rng(1)
    sims, freqs = [], []
    for k in range(n_files):
        q, fq = os.path.join(d, f"run{k}.tsv"), os.path.join(d, f"run{k}_freqs.txt")
        if k < 32:
            a = wl.sim_batch_member(k)
            a[:, 0] = np.round(a[:, 0], 6)
            with open(q, "w") as f:
                f.writelines(f"{x!r}\t{y!r}\n" for x, y in a.tolist())
            np.savetxt(fq, np.round(rng.beta(0.3, 0.3, 5000), 4), fmt="%.4f")
        else:
            shutil.copy(sims[k % 32], q)
            shutil.copy(freqs[k % 32], fq)
        sims.append(q)
        freqs.append(fq)
    observed = [0.0, 0.3, 0.2, 0.5]
    print(f"{n_files} simulations of 100 000 rows + 5 000 gene frequencies each; {os.cpu_count()} host cores", flush=True)
    out = {}
    for parse in ("host", "device"):
        simulator.process_results_batch(sims[:64], freqs[:64], T, 0.25, 0.95, 0.05, observed=observed, parse=parse)   # warm-up
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            out[parse] = simulator.process_results_batch(sims, freqs, T, 0.25, 0.95, 0.05, observed=observed, parse=parse)
            best = min(best, time.perf_counter() - t0)
        print(f"process_results_batch, files -> (summary, d, log d), tables parsed on the {parse}: "
              f"{n_files / best:.0f} simulations/s ({best:.3f} s)", flush=True)
    for a, b in zip(out["host"], out["device"]):
        assert np.array_equal(a, b)
    print("bit-identical between the two pipelines", flush=True)
finally:
    shutil.rmtree(d, ignore_errors=True)
