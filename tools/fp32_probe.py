"""GPU probe: candidate-loop ceilings and the bookkeeping of the fp32 culled kernel on the bench workload.
usage: [KDEJS_FP32_THR=x] python tools/fp32_probe.py"""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import poppunk_mod_b200 as pk
from oracle import workloads as wl

nat = pk._native
if os.environ.get("PROBE_LIB"):                       # A/B against another build of the library
    nat.LIB_PATH = os.path.abspath(os.environ["PROBE_LIB"])
L = nat.lib(); dev = 0
nat.check(L.kdejs_init(dev))
if os.environ.get("PROBE_LOOPS", "1") == "1":
    for which, name in ((0, "fp64 loop"), (1, "fp32 loop"), (2, "fp32 loop + fp64 fold every 16")):
        v = ctypes.c_double()
        nat.check(L.kdejs_loop_peak(dev, which, ctypes.byref(v)))
        cyc = 148 * 4 * 1.965e9 / (v.value / 512.0)          # SMSP cycles per warp trip (32 candidates x 16 nodes)
        print(f"loop peak {name}: {v.value:.3e} pair tests/s = {cyc:.1f} cycles per warp trip per SMSP", flush=True)
B = 64
T = wl.example_target()
sims = [wl.sim_batch_member(s) for s in range(B)]
dT = torch.from_numpy(T).cuda(); dS = [torch.from_numpy(s).cuda() for s in sims]
out = torch.zeros(B, dtype=torch.float64, device="cuda")
ptrs = (ctypes.c_void_p * B)(*[t.data_ptr() for t in dS]); cnt = (ctypes.c_int64 * B)(*([100000] * B))
st = torch.cuda.Stream()
R = int(os.environ.get("PROBE_R", 256))
def run(algo):
    nat.check(L.kdejs_discrepancy_batch_dev(dev, dT.data_ptr(), 100000, ptrs, cnt, B, R, 0.03, 0.25, 0.0, 0, 0,
                                            nat.ALGOS[algo], out.data_ptr(), None, st.cuda_stream))
for algo in ("binned", "binned32"):
    for _ in range(3): run(algo)
    torch.cuda.synchronize()
    nat.check(L.kdejs_profile_reset(dev)); nat.check(L.kdejs_profile_enable(dev, 1))
    s5 = (ctypes.c_double * 5)(); nat.check(L.kdejs_fp32_stats(dev, s5, 1))
    for _ in range(10): run(algo)
    torch.cuda.synchronize()
    pm, pn = ctypes.c_double(), ctypes.c_int64()
    prof = {}
    for k, nm in enumerate(nat.KERNEL_NAMES):
        nat.check(L.kdejs_profile_read(dev, k, ctypes.byref(pm), ctypes.byref(pn))); prof[nm] = round(1e3 * pm.value / (10 * B), 3)
    nat.check(L.kdejs_fp32_stats(dev, s5, 1)); nat.check(L.kdejs_profile_enable(dev, 0))
    print(algo, "R", R, "thr", os.environ.get("KDEJS_FP32_THR"), "us/eval", prof, flush=True)
    if algo == "binned32":
        t = [x / (10 * B) for x in s5]
        print(f"  per eval: tiles {t[0]:.0f}, second walks {t[1]:.0f} ({100*t[1]/max(t[0],1):.1f} %), warp trips first {t[2]:.0f}, "
              f"second {t[3]:.0f} ({100*t[3]/max(t[2],1):.1f} %), re-queued {t[4]:.1f}", flush=True)
    js = out.cpu().numpy().copy()
    if algo == "binned": exact = js
    else: print("  max rel gap to fp64:", float(np.max(np.abs(js - exact) / exact)), flush=True)
