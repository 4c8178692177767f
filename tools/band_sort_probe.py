"""GPU probe: the kernels of ONE large evaluation on one device (config C5's geometry, R = 1024, 128 x 128 cells).
usage: python tools/band_sort_probe.py [points per set]   (1.25 M = the local sort of one of eight bands of C5)"""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import poppunk_mod_b200 as pk
from oracle import workloads as wl

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_250_000
nat = pk._native; L = nat.lib(); dev = 0
a, b = wl.c5_inputs(n, seed=5)
dA, dB = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
out = torch.zeros(1, dtype=torch.float64, device="cuda")
ptrs = (ctypes.c_void_p * 1)(dB.data_ptr()); cnt = (ctypes.c_int64 * 1)(n)
st = torch.cuda.Stream()
def run():
    nat.check(L.kdejs_discrepancy_batch_dev(dev, dA.data_ptr(), n, ptrs, cnt, 1, 1024, 0.03, 0.25, 0.0, 0, 0,
                                            nat.ALGOS["binned"], out.data_ptr(), None, st.cuda_stream))
for _ in range(3): run()
torch.cuda.synchronize()
nat.check(L.kdejs_profile_reset(dev)); nat.check(L.kdejs_profile_enable(dev, 1))
t0 = time.perf_counter()
for _ in range(5): run()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
pm, pn = ctypes.c_double(), ctypes.c_int64()
prof = {}
for k, nm in enumerate(nat.KERNEL_NAMES):
    nat.check(L.kdejs_profile_read(dev, k, ctypes.byref(pm), ctypes.byref(pn))); prof[nm] = round(pm.value / 5, 4)
nat.check(L.kdejs_profile_enable(dev, 0))
print(f"n={n} R=1024: {1e3*dt:.3f} ms per evaluation (with timers); kernel ms {prof}; JS {float(out[0]):.12f}", flush=True)
