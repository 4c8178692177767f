"""GPU: band mode (one evaluation with the points sharded over ranks, kdejs_band_*) on ONE device: the ranks of a
virtual world run one after the other through the same C-ABI steps distributed.band_sharded_discrepancy issues, the
collectives in between are plain concatenations.  (The NCCL path itself is exercised by tools/mgpu_check.py and
bench.py --gpus N on real multi-GPU boxes.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pk():
    import poppunk_mod_b200 as pkg

    assert pkg._native.device_count() > 0
    return pkg


def virtual_band(pk, a, b, world, gamma, eps, log=False, resolution=100, bandwidth=0.03, algo="binned", cuts=None):
    """band_sharded_discrepancy's sequence for `world` virtual ranks on device 0."""
    import torch

    from poppunk_mod_b200 import distributed as D

    common = pk.KDE_distance._common(resolution, bandwidth, gamma, eps, log, "epanechnikov", algo)
    ops = D._GpuBandOps(0, common)
    ops.peer = False          # the virtual ranks share one device, hence one exchange buffer: keep the staged path
    with ops.run():
        return _virtual_band(torch, D, ops, a, b, world, resolution, bandwidth, cuts)


def _virtual_band(torch, D, ops, a, b, world, resolution, bandwidth, cuts):
    sl = lambda x, r: x[len(x) * r // world: len(x) * (r + 1) // world]
    A = [ops.to_device(np.ascontiguousarray(sl(a, r)), "a") for r in range(world)]
    B = [ops.to_device(np.ascontiguousarray(sl(b, r)), "b") for r in range(world)]
    stats_all = torch.cat([ops.minmax(A[r], B[r]) for r in range(world)])
    srt, blocks = zip(*[ops.sort(A[r], B[r], stats_all, world) for r in range(world)])
    blocks_all = torch.cat(blocks)
    colwork, rs, use_m = ops.plan(blocks_all, world)
    cuts = cuts or D.band_cuts(colwork, world)
    rows = [D.band_cell_rows(resolution, bandwidth, cuts[d], cuts[d + 1]) for d in range(world)]
    n_tot = rs[:, :, -1].sum(axis=0)

    def stage_for(d, st):
        lo, hi = rows[d]
        if hi < lo:
            return ops.empty((0, 2))
        parts = []
        for s_ in range(world):
            base = 0 if st == 0 else A[s_].shape[0]
            parts.append(srt[s_][base + rs[s_, st, lo]: base + rs[s_, st, hi + 1]])
        return torch.cat(parts).contiguous()

    def phase1(d):
        return ops.phase1(stage_for(d, 0), stage_for(d, 1), n_tot[0], n_tot[1], blocks_all, stats_all, world,
                          (cuts[d], cuts[d + 1]), rows[d], use_m)

    norms_all = torch.cat([phase1(d) for d in range(world)])
    lrs = []
    for d in range(world):
        phase1(d)                               # phase 2 continues from the state phase 1 left on the device
        lrs.append(ops.to_host(ops.phase2(norms_all, world)))
    parts = np.array(lrs)
    assert not parts[:, 2].any()
    return np.float64(np.sqrt(max((parts[:, 0].sum() + parts[:, 1].sum()) / 2.0, 0.0))), cuts, rows, use_m


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("algo", ["binned", "binned32"])
def test_band_mode_equals_single_gpu(pk, world, algo):
    from oracle import kde_oracle as ko
    from oracle import workloads as wl

    a, b = wl.synthetic_target(60000), wl.sim_S(3, 1500, 0.33, n=45000)
    R = 130
    full = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=R, algo=algo)
    js, cuts, rows, _ = virtual_band(pk, a, b, world, 0.25, 0.0, resolution=R, algo=algo)
    assert cuts[0] == 0 and cuts[-1] == (R + 3) // 4 and all(x <= y for x, y in zip(cuts, cuts[1:]))
    assert js == pytest.approx(full, rel=1e-12)
    want = ko.discrepancy(a, b, 0.25, 0.0, resolution=R)
    assert abs(js - want) <= (1e-10 if algo == "binned" else 2e-7) * want


def test_band_mode_fine_grid_with_moment_cells_and_empty_bands(pk, monkeypatch):
    """R = 1024 (128 x 128 cells, moment cells), eight bands of a concentrated cloud: some bands are thin, the
    explicit cuts below also leave two ranks without any tile column."""
    from oracle import workloads as wl

    monkeypatch.setenv("KDEJS_MOMENTS", "1")       # 400 k points are below the points-per-cell rule: force the path
    a, b = wl.c5_inputs(400_000, seed=3)
    full = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=1024)
    js, cuts, rows, use_m = virtual_band(pk, a, b, 8, 0.25, 0.0, resolution=1024)
    assert use_m == 1
    assert js == pytest.approx(full, rel=1e-12)
    forced = [0, 0, 40, 41, 120, 200, 256, 256, 256]
    js2, _, rows2, _ = virtual_band(pk, a, b, 8, 0.25, 0.0, resolution=1024, cuts=forced)
    assert rows2[0][1] < rows2[0][0] and rows2[7][1] < rows2[7][0]
    assert js2 == pytest.approx(full, rel=1e-12)


def test_band_mode_log_nan_and_ragged_slices(pk):
    from oracle import kde_oracle as ko
    from oracle import workloads as wl

    a, b = wl.ragged_inputs()
    a = np.abs(a) + 1e-3
    b = np.abs(b) + 1e-3
    for gamma, eps, log in ((0.25, 0.0, False), (1e-12, 1e-12, True), (1.0, 1e-12, False)):
        js, *_ = virtual_band(pk, a, b, 3, gamma, eps, log)
        want = ko.discrepancy(a, b, gamma, eps, log)
        assert abs(js - want) <= 1e-10 * want + 1e-12


def test_band_sharded_discrepancy_world_one(pk):
    """The public entry without a process group (world = 1): host arrays and device tensors."""
    import torch

    from oracle import workloads as wl
    from poppunk_mod_b200 import distributed as D

    a, b = wl.synthetic_target(50000), wl.sim_S(4, 2000, 0.3, n=50000)
    full = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=256)
    info = {}
    js = D.band_sharded_discrepancy(a, b, 0.25, 0.0, resolution=256, info=info)
    assert js == pytest.approx(full, rel=1e-12) and info["tile_columns"] == (0, 64)
    js_d = D.row_sharded_discrepancy(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 0.25, 0.0,
                                     resolution=256, sliced=True)
    assert js_d == js
    bad = b.copy()
    bad[7, 0] = np.inf
    with pytest.raises(ValueError, match="infinity"):
        D.band_sharded_discrepancy(a, bad, 0.25, 0.0, resolution=256)


@pytest.mark.parametrize("algo", ["binned", "binned32"])
def test_merge_from_the_exchange_buffer_world_one(pk, algo):
    """kdejs_band_peer_*: the merge kernel reads the sorted points out of the rank's exported exchange buffer (with
    one rank: its own).  Same answer as the evaluation on one GPU; the buffer grows and is exported again on demand."""
    import ctypes

    import torch

    from oracle import workloads as wl
    from poppunk_mod_b200 import distributed as D

    nat, L = pk._native, pk._native.lib()
    for n in (30_000, 90_000):                       # the second size makes the buffer grow: a new handle
        a, b = wl.synthetic_target(n), wl.sim_S(4, 1700, 0.3, n=n // 2)
        R = 200
        common = pk.KDE_distance._common(R, 0.03, 0.25, 0.0, False, "epanechnikov", algo)
        ops = D._GpuBandOps(0, common)
        with ops.run():
            A, B = ops.to_device(a, "a"), ops.to_device(b, "b")
            stats_all = ops.minmax(A, B)
            srt, handle = ops._peer_buffer(len(a) + len(b))
            block = ops.empty(ops.block_u32, "i4")
            block[-D.HANDLE_WORDS:] = handle
            nat.check(L.kdejs_band_sort_dev(0, A.data_ptr(), len(a), B.data_ptr(), len(b), *common, stats_all.data_ptr(), 1,
                                            srt.data_ptr(), block.data_ptr(), ops._stream()))
            colwork, rs, use_m = ops.plan(block, 1)
            assert ops.handles.any()
            nat.check(L.kdejs_band_peer_open(0, ops.handles.ctypes.data, 1, 0))
            NT = ops.NT
            rows = D.band_cell_rows(R, 0.03, 0, NT)
            norm = ops.phase1_peer(len(a), len(b), len(a), len(b), block, stats_all, 1, (0, NT), rows, use_m)
            lr = ops.to_host(ops.phase2(norm, 1))
        js = np.sqrt(max((lr[0] + lr[1]) / 2.0, 0.0))
        assert lr[2] == 0
        assert js == pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=R, algo=algo)
    # a handle that is not this rank's own is refused
    bad = np.array(ops.handles, copy=True)
    bad[0, 0] ^= 1
    with pytest.raises(ValueError):
        nat.check(L.kdejs_band_peer_open(0, bad.ctypes.data, 1, 0))
