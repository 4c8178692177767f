"""The device table parser alone: a wave of 16-64 tables held as text in page-locked memory through
kdejs_discrepancy_batch_text (for ncu captures of k_tsv_count / k_tsv_parse).  usage: python tools/tsv_probe.py [tables]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import poppunk_mod_b200 as pk
from oracle import workloads as wl

nat = pk._native
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = wl.example_target()
texts = []
for s in range(8):
    a = wl.sim_batch_member(s)
    a[:, 0] = np.round(a[:, 0], 6)
    texts.append("".join(f"{x!r}\t{y!r}\n" for x, y in a.tolist()).encode())
slot = max(len(t) for t in texts) + 4096
arena = nat.TextArena(B, slot)
nb = np.zeros(B, dtype=np.int64)
for k in range(B):
    t = texts[k % 8]
    arena.array[k, :len(t)] = np.frombuffer(t, dtype=np.uint8)
    nb[k] = len(t)
for wave, tail in [(int(x.split("/")[0]), x.split("/")[1] if "/" in x else "") for x in os.environ.get("TEXT_WAVES", "16").split(",")]:
    os.environ["KDEJS_TEXT_WAVE"] = str(wave)
    os.environ.pop("KDEJS_TEXT_TAIL", None)
    if tail:
        os.environ["KDEJS_TEXT_TAIL"] = tail
    best = 1e9
    for rep in range(6):
        t0 = time.perf_counter()
        js, rows, status, bnd = pk.KDE_JS_divergence_batch_text(T, arena.array, 0.25, 0.0, nbytes=nb, skip_rows=0, bounds=True)
        best = min(best, time.perf_counter() - t0)
    assert (status == 0).all() and (rows == 100000).all()
    print(f"{B} tables of {nb[0] / 1e6:.2f} MB, waves of {wave} (last wave: {tail or "default"}): {1e3 * best:.2f} ms per call = {B / best:.0f} tables/s; js[0] = {js[0]!r}")
