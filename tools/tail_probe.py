"""Host batch path, 64 evaluations per call (the bench's e2e step): first / middle / last wave sizes, A/B in one process.
usage: python tools/tail_probe.py  (on the GPU box)"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import poppunk_mod_b200 as pk
from oracle import workloads as wl

nat = pk._native
pool = nat.PinnedPool()
T = pool.empty((100000, 2)); T[:] = wl.example_target()
B = 64
arena = pool.empty((2 * B, 100000, 2))      # one page-locked (2B,N,2) array; `sims` are its rows as a list
sims = [arena[i] for i in range(2 * B)]
for s, buf in enumerate(sims):
    buf[:] = wl.sim_batch_member(s)
AS_3D = False
kw = dict(gamma=0.25, eps=0.0, resolution=256)


def run(steps=12):
    src = arena if AS_3D else sims
    for s in range(2):
        pk.KDE_JS_divergence_batch(T, src[(s % 2) * B:(s % 2) * B + B], **kw)
    t0 = time.perf_counter()
    for s in range(steps):
        r = pk.KDE_JS_divergence_batch(T, src[(s % 2) * B:(s % 2) * B + B], **kw)
    return steps * B / (time.perf_counter() - t0), r


ref = None
configs = [dict(), dict(KDEJS_TAIL_WAVE="6"), dict(KDEJS_TAIL_WAVE="4"), dict(KDEJS_TAIL_WAVE="8"),
           dict(KDEJS_TAIL_WAVE="6", KDEJS_FIRST_WAVE="2"), dict(KDEJS_TAIL_WAVE="6", KDEJS_HOST_WAVE="12"),
           dict(KDEJS_TAIL_WAVE="6", KDEJS_HOST_WAVE="20"), dict(KDEJS_FIRST_WAVE="2"), dict(KDEJS_FIRST_WAVE="8")]
if len(sys.argv) > 1:
    configs = [dict(kv.split("=") for kv in a.split(";") if kv) for a in sys.argv[1:]]
for rep in range(3):
    for cfg in configs:
        cfg = dict(cfg)
        for k in ("KDEJS_TAIL_WAVE", "KDEJS_FIRST_WAVE", "KDEJS_HOST_WAVE", "KDEJS_WAVE_PLAN"):
            os.environ.pop(k, None)
        AS_3D = cfg.pop("3D", "0") == "1"
        os.environ.update(cfg)
        v, r = run()
        cfg = dict(cfg, **{"3D": int(AS_3D)})
        if ref is None:
            ref = r.copy()
        assert np.array_equal(r, ref), "results changed with the wave sizes"
        print(f"rep {rep} {cfg}: {v:.0f} evals/s", flush=True)
