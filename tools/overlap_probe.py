"""GPU probe: device-resident throughput with ONE caller stream vs TWO alternating caller streams (the library gives
each stream its own scratch slot, so the sort of one wave can run beside the density kernel of the other).
usage: [KDEJS_DENS_BLOCKS=n] python tools/overlap_probe.py [algo]"""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import poppunk_mod_b200 as pk
from oracle import workloads as wl

algo = sys.argv[1] if len(sys.argv) > 1 else "binned32"
nat = pk._native; L = nat.lib(); dev = 0
nat.check(L.kdejs_init(dev))
B, POOL = 64, 192
T = wl.example_target()
sims = [wl.sim_batch_member(s) for s in range(POOL)]
dT = torch.from_numpy(T).cuda(); dS = [torch.from_numpy(s).cuda() for s in sims]
outs = [torch.zeros(B, dtype=torch.float64, device="cuda") for _ in range(2)]
cnt = (ctypes.c_int64 * B)(*([100000] * B))
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def run(step, k):
    first = (step * B) % POOL
    ptrs = (ctypes.c_void_p * B)(*[dS[(first + i) % POOL].data_ptr() for i in range(B)])
    nat.check(L.kdejs_discrepancy_batch_dev(dev, dT.data_ptr(), 100000, ptrs, cnt, B, 256, 0.03, 0.25, 0.0, 0, 0,
                                            nat.ALGOS[algo], outs[k].data_ptr(), None, streams[k].cuda_stream))
for nstreams in (1, 2):
    for s in range(6): run(s, s % nstreams)
    torch.cuda.synchronize()
    steps = 200
    t0 = time.perf_counter()
    for s in range(steps): run(s, s % nstreams)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{algo} dens_blocks={os.environ.get('KDEJS_DENS_BLOCKS')} streams={nstreams}: {steps*B/dt:.0f} evals/s  ({1e6*dt/(steps*B):.2f} us/eval)", flush=True)
run(0, 0); run(0, 1); torch.cuda.synchronize()
assert np.array_equal(outs[0].cpu().numpy(), outs[1].cpu().numpy())
