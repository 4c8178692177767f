"""Device-resident waves of 1 ... 64 evaluations (bench workload): time per wave and per kernel -- the fixed cost a wave
pays whatever its size.  usage: python tools/wave_probe.py  (on the GPU box)"""
import ctypes, os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
import poppunk_mod_b200 as pk

nat = pk._native; L = nat.lib(); dev = 0
torch.cuda.set_device(dev)
target, sims = bench.make_inputs(0, bench.POOL)
db = bench.DeviceBench(torch, nat, dev, target, sims)
algo = os.environ.get("ALGO", "binned32")
splits = [tuple(a.split(",")) for a in sys.argv[1:]] or [("", "")]
for split, batch in [(sp, b) for sp in splits for b in [int(x) for x in os.environ.get("WAVES", "1,2,4,8,16,64").split(",")]]:
    for k, v in zip(("KDEJS_SPLIT_LOG2", "KDEJS_PART_LOG2"), split):
        os.environ.pop(k, None)
        if v:
            os.environ[k] = v
    steps = max(20, 640 // batch)
    for s in range(5):
        db.step(s, 256, algo, batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(db.stream)
    for s in range(steps):
        db.step(5 + s, 256, algo, batch)
    e1.record(db.stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    nat.check(L.kdejs_profile_reset(dev)); nat.check(L.kdejs_profile_enable(dev, 1))
    for s in range(20):
        db.step(s, 256, algo, batch)
    torch.cuda.synchronize()
    pm, pn = ctypes.c_double(), ctypes.c_int64(); prof = {}
    for k, nm in enumerate(nat.KERNEL_NAMES):
        nat.check(L.kdejs_profile_read(dev, k, ctypes.byref(pm), ctypes.byref(pn)))
        if pn.value:
            prof[nm] = round(1e3 * pm.value / 20, 1)
    nat.check(L.kdejs_profile_enable(dev, 0))
    print(f"split {split} wave of {batch:2d}: {1e3 * ms:7.1f} us ({1e3 * ms / batch:5.1f} per evaluation); kernel us {prof} sum {sum(prof.values()):.0f}", flush=True)
