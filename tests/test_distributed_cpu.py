"""CPU, world_size 2 over gloo: the host-side sharding logic of poppunk-mod_b200/distributed.py.
The per-rank compute is injected (the CPU oracle stands in for the GPU kernels, which cannot run here);
what is under test is item dealing, row ranges, gather order and the two-phase reduction."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import poppunk_mod_b200 as pk
    from oracle import kde_oracle as ko, workloads as wl
    from poppunk_mod_b200 import distributed as D

    obs = wl.synthetic_target(3000)
    sims = [wl.sim_batch_member(s, n=1000 + 300 * s) for s in range(7)]

    def compute(o, ss, gamma, eps, log, **kw):
        return np.array([ko.discrepancy(o, s, gamma, eps, log) for s in ss])

    batch = D.batch_discrepancy_sharded(obs, sims, 0.25, 0.0, False, compute=compute)

    R = 50
    a, b = obs, sims[3]
    _, z1, z2, d1, d2 = ko.discrepancy(a, b, 0.25, 0.0, resolution=R, return_grids=True)
    P = [np.sqrt(np.sqrt(d)) for d in (d1, d2)]        # **0.25 + eps(0), unnormalised
    state = {}

    def phase1(df1, df2, lo, hi):
        state["rows"] = (lo, hi)
        return np.array([P[0][lo * R:hi * R].sum(), P[1][lo * R:hi * R].sum()])

    def phase2(tot):
        lo, hi = state["rows"]
        p, q = P[0][lo * R:hi * R] / tot[0], P[1][lo * R:hi * R] / tot[1]
        m = (p + q) / 2
        with np.errstate(divide="ignore", invalid="ignore"):
            l = np.where(p > 0, p * np.log(p / m), 0.0).sum()
            r = np.where(q > 0, q * np.log(q / m), 0.0).sum()
        return np.array([l, r])

    js = D.row_sharded_discrepancy(a, b, 0.25, 0.0, resolution=R, phase1=phase1, phase2=phase2)
    # sliced inputs: each rank holds a ragged contiguous slice; the all-gather must rebuild the table
    cut_a, cut_b = (0, 1700, len(a)), (0, 901, len(b))
    seen = {}

    def phase1_sliced(df1, df2, lo, hi):
        seen["ok"] = bool(np.array_equal(df1, a) and np.array_equal(df2, b))
        return phase1(df1, df2, lo, hi)

    js_sliced = D.row_sharded_discrepancy(a[cut_a[rank]:cut_a[rank + 1]], b[cut_b[rank]:cut_b[rank + 1]], 0.25, 0.0,
                                          resolution=R, sliced=True, phase1=phase1_sliced, phase2=phase2)
    assert seen["ok"] and js_sliced == js
    q.put((rank, batch, float(js), float(ko.discrepancy(a, b, 0.25, 0.0, resolution=R)),
           [float(ko.discrepancy(obs, s, 0.25, 0.0)) for s in sims]))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    (r0, b0, js0, want_js, want_batch), (r1, b1, js1, _, _) = res
    np.testing.assert_array_equal(b0, b1)                 # every rank holds the full, identical result
    np.testing.assert_allclose(b0, want_batch, rtol=1e-13)
    assert js0 == js1                                     # rank-ordered sums: bit-identical across ranks
    assert js0 == pytest.approx(want_js, rel=1e-12)


def test_row_ranges_and_item_dealing():
    import poppunk_mod_b200  # noqa: F401
    from poppunk_mod_b200 import distributed as D

    for R in (1, 3, 4, 100, 130, 256, 1024):
        for world in (1, 2, 3, 8):
            rr = D.row_ranges(R, world)
            assert rr[0][0] == 0 and rr[-1][1] == R and len(rr) == world
            for (lo, hi), (lo2, _) in zip(rr[:-1], rr[1:]):
                assert hi == lo2 and lo <= hi
            for lo, hi in rr:
                assert lo % 4 == 0 and (hi % 4 == 0 or hi == R)
    assert sorted(sum((D.shard_items(10, r, 3) for r in range(3)), [])) == list(range(10))
