"""GPU probe: issue-rate peaks, per-kernel timings of the headline workload, quick parity."""
import ctypes
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import poppunk_mod_b200 as pk
from oracle import kde_oracle as ko, workloads as wl

nat = pk._native
L = nat.lib()
nat.check(L.kdejs_init(0))
for which, name in ((0, "FFMA fp32"), (1, "DFMA fp64"), (2, "FFMA2 f32x2")):
    tf, mhz = ctypes.c_double(), ctypes.c_double()
    nat.check(L.kdejs_peak_flops(0, which, ctypes.byref(tf), ctypes.byref(mhz)))
    print(f"peak {name}: {tf.value:.2f} TFLOP/s (max clock {mhz.value:.0f} MHz)", flush=True)

T = wl.synthetic_target(100_000)
S = wl.sim_S(254, 2500, 0.38)
for R in (100, 256):
    t0 = time.time(); want = ko.discrepancy(T, S, 0.25, 0.0, resolution=R); t_or = time.time() - t0
    got = pk.KDE_JS_divergence(T, S, gamma=0.25, eps=0.0, resolution=R)
    t0 = time.time()
    for _ in range(10):
        got = pk.KDE_JS_divergence(T, S, gamma=0.25, eps=0.0, resolution=R)
    dt = (time.time() - t0) / 10
    print(f"R={R}: gpu={got!r} oracle={want!r} rel={abs(got-want)/want:.2e}  single-call {dt*1e3:.3f} ms (oracle {t_or*1e3:.0f} ms)", flush=True)

sims = [wl.sim_batch_member(s) for s in range(64)]
for R in (100, 256):
    nat.check(L.kdejs_profile_reset(0)); nat.check(L.kdejs_profile_enable(0, 1))
    w = ctypes.c_double(); nat.check(L.kdejs_work_counters(0, ctypes.byref(w), 1))
    out = pk.KDE_JS_divergence_batch(T, sims, gamma=0.25, eps=0.0, resolution=R)
    t0 = time.time(); out = pk.KDE_JS_divergence_batch(T, sims, gamma=0.25, eps=0.0, resolution=R); dt = time.time() - t0
    nat.check(L.kdejs_work_counters(0, ctypes.byref(w), 1))
    print(f"batch64 R={R}: {dt*1e3:.2f} ms host-to-host -> {64/dt:.0f} evals/s ; pair tests/eval {w.value/128:.3e}")
    for k, nm in enumerate(nat.KERNEL_NAMES):
        ms, n = ctypes.c_double(), ctypes.c_int64()
        nat.check(L.kdejs_profile_read(0, k, ctypes.byref(ms), ctypes.byref(n)))
        print(f"   {nm:12s} {ms.value:9.3f} ms over {n.value} launches", flush=True)
    nat.check(L.kdejs_profile_enable(0, 0))
    chk = [ko.discrepancy(T, sims[i], 0.25, 0.0, resolution=R) for i in (0, 63)]
    print("   parity", abs(out[0]-chk[0])/chk[0], abs(out[63]-chk[1])/chk[1])
