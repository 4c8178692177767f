"""CPU, world_size 2 and 3 over gloo: the exchange logic of distributed.band_sharded_discrepancy (which cell rows a
band needs, what every rank sends to whom, where received slices land, rank-ordered folds).  The device steps are
replaced by a numpy stand-in with the same interface (the CPU oracle computes the band's densities); the host-only
C-ABI helpers (kdejs_band_geometry / _cuts / _cell_rows) are the real ones."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


class NumpyBandOps:
    """Same methods as distributed._GpuBandOps, torch CPU tensors in and out."""

    def __init__(self, common):
        import torch

        import poppunk_mod_b200 as pk

        self.torch, self.common = torch, common
        self.R, self.h, self.gamma, self.eps, self.log = common[0], common[1], common[2], common[3], common[4]
        geo = (ctypes.c_int * 4)()
        pk._native.check(pk._native.lib().kdejs_band_geometry(self.R, self.h, 0, geo))
        self.C, self.NT, self.block_u32 = geo[0], geo[1], geo[2]
        self.state = {}

    def run(self):
        import contextlib

        return contextlib.nullcontext()

    def to_device(self, x, name):
        return self.torch.from_numpy(np.array(x, dtype=np.float64, order="C"))

    def empty(self, shape, dtype="f8"):
        return self.torch.zeros(shape, dtype={"f8": self.torch.float64, "i4": self.torch.int32}[dtype])

    def _pre(self, t):
        x = t.numpy()
        return np.log(x + self.eps) if self.log else x

    def minmax(self, a, b):
        out = []
        for t in (a, b):
            x = self._pre(t)
            out += [np.nanmin(x[:, 0]), np.nanmin(x[:, 1]), np.nanmax(x[:, 0]), np.nanmax(x[:, 1]), 0.0]
        return self.torch.tensor(out, dtype=self.torch.float64)

    def _scaler(self, stats_all, world):
        s = stats_all.numpy().reshape(world, 2, 5)
        mn = np.minimum(s[:, :, 0:2].min(axis=(0, 1)), np.inf)
        mx = s[:, :, 2:4].max(axis=(0, 1))
        rng = mx - mn
        rng[rng < 10 * np.finfo(float).eps] = 1.0
        scale = 1.0 / rng
        return scale, -mn * scale

    def sort(self, a, b, stats_all, world):
        scale, min_ = self._scaler(stats_all, world)
        C, nc = self.C, self.C * self.C
        srt, cs, rs = [], [], []
        for t in (a, b):
            X = self._pre(t) * scale + min_
            X[np.isnan(X)] = 0.0
            cell = (np.clip(np.floor(X[:, 1] * C), 0, C - 1) * C + np.clip(np.floor(X[:, 0] * C), 0, C - 1)).astype(np.int64)
            order = np.argsort(cell, kind="stable")
            srt.append(X[order])
            start = np.concatenate([[0], np.cumsum(np.bincount(cell, minlength=nc))])
            cs.append(start)
            rs.append(start[::C][:C + 1])
        block = np.concatenate(cs + rs + [np.zeros(16)]).astype(np.int32)    # 16 words: no peer-memory handle
        assert len(block) == self.block_u32
        return self.torch.from_numpy(np.vstack(srt)), self.torch.from_numpy(block)

    def plan(self, blocks_all, world):
        C, nc = self.C, self.C * self.C
        blk = blocks_all.numpy().reshape(world, -1)
        rowstart = blk[:, 2 * (nc + 1):-16].reshape(world, 2, C + 1).astype(np.int64)
        per_row = np.diff(rowstart, axis=2).sum(axis=(0, 1)).astype(np.float64)           # points per cell row
        colwork = 1.0 + np.interp(np.linspace(0, C - 1, self.NT), np.arange(C), per_row)  # any positive proxy will do
        return colwork, rowstart, 0

    def phase1(self, stage_a, stage_b, na_tot, nb_tot, blocks_all, stats_all, world, ty, cy, use_m):
        from oracle import kde_oracle as ko

        R = self.R
        iy = np.arange(4 * ty[0], min(4 * ty[1], R))
        P = []
        for stage, ntot in ((stage_a, na_tot), (stage_b, nb_tot)):
            pts = stage.numpy()
            if len(pts):
                d = ko.kde_density(pts, R, self.h).reshape(R, R)[:, iy] * (len(pts) / float(ntot))
            else:
                d = np.zeros((R, len(iy)))
            P.append(np.power(d, self.gamma) + self.eps)
        self.state["P"] = P
        return self.torch.tensor([P[0].sum(), P[1].sum()], dtype=self.torch.float64)

    def phase2(self, norms_all, world):
        tot = norms_all.numpy().reshape(world, 2).sum(axis=0)
        p, q = self.state["P"][0] / tot[0], self.state["P"][1] / tot[1]
        m = (p + q) / 2
        with np.errstate(divide="ignore", invalid="ignore"):
            node = np.where(p > 0, p * np.log(p / m), 0.0) + np.where(q > 0, q * np.log(q / m), 0.0)
        return self.torch.tensor([node.sum(), 0.0, 0.0], dtype=self.torch.float64)

    def to_host(self, t):
        return t.numpy()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import poppunk_mod_b200 as pk
    from oracle import kde_oracle as ko, workloads as wl
    from poppunk_mod_b200 import distributed as D
    from test_band_cpu import NumpyBandOps

    a, b = wl.synthetic_target(4000), wl.sim_S(3, 1500, 0.33, n=3100)
    out = []
    for R, gamma, eps, log in ((64, 0.25, 0.0, False), (100, 1.0, 1e-12, False), (37, 0.5, 1e-12, True)):
        common = pk.KDE_distance._common(R, 0.03, gamma, eps, log, "epanechnikov", "binned")
        cuts_a = [len(a) * r // world + (7 if 0 < r < world else 0) for r in range(world + 1)]      # ragged slices
        cuts_b = [len(b) * r // world for r in range(world + 1)]
        info = {}
        js = D.band_sharded_discrepancy(a[cuts_a[rank]:cuts_a[rank + 1]], b[cuts_b[rank]:cuts_b[rank + 1]], gamma, eps,
                                        log, resolution=R, ops=NumpyBandOps(common), info=info)
        out.append((float(js), float(ko.discrepancy(a, b, gamma, eps, log, resolution=R)), info))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_exchange_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for case in range(3):
        vals = [res[r][1][case][0] for r in range(world)]
        want = res[0][1][case][1]
        assert all(v == vals[0] for v in vals)                 # rank-ordered folds: identical on every rank
        assert vals[0] == pytest.approx(want, rel=1e-11)
        cols = [res[r][1][case][2]["tile_columns"] for r in range(world)]
        assert cols[0][0] == 0 and all(cols[r][1] == cols[r + 1][0] for r in range(world - 1))


def test_band_host_helpers():
    import poppunk_mod_b200 as pk
    from poppunk_mod_b200 import distributed as D

    geo = (ctypes.c_int * 4)()
    pk._native.check(pk._native.lib().kdejs_band_geometry(1024, 0.03, 0, geo))
    assert list(geo)[:3] == [128, 256, 2 * (128 * 128 + 1) + 2 * 129 + 16]      # + the peer-memory handle words
    cw = np.ones(256)
    cw[100:130] = 50.0
    for world in (1, 2, 5, 8):
        cuts = D.band_cuts(cw, world)
        assert cuts[0] == 0 and cuts[-1] == 256 and all(x <= y for x, y in zip(cuts, cuts[1:]))
        work = [cw[cuts[r]:cuts[r + 1]].sum() for r in range(world)]
        assert max(work) <= cw.sum() / world + 50.0           # no band exceeds its share by more than one column
    assert D.band_cell_rows(1024, 0.03, 0, 256) == (0, 127)
    lo, hi = D.band_cell_rows(1024, 0.03, 100, 110)
    # tile columns 100..109 cover y in [400/1023, 439/1023]; with h = 0.03 the rows are those of y -/+ h
    assert lo == int(np.floor((400 / 1023 - 0.03) * 128)) and hi == int(np.floor((439 / 1023 + 0.03) * 128))
    assert D.band_cell_rows(1024, 0.03, 7, 7) == (0, -1)      # empty band
    with pytest.raises(ValueError):
        D.band_cell_rows(1024, 0.03, 5, 300)
