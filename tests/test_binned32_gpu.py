"""GPU: the culled sum on the FP32 pipe (algo="binned32", k_density_binned32) against the exact oracle, the golden
vectors and the fp64 culled path.  Bars are SURVEY.md 8c(i)'s for an fp32 kernel: JS within 1e-5 relative (observed
~1e-8), per-node density within 1e-6 z + 1e-7 z_max -- and, beyond that protocol, exact zeros stay exact zeros."""
import ctypes

import numpy as np
import pytest

from conftest import case_kwargs, golden_inputs
from test_parity_gpu import ALL_CASES

pytestmark = pytest.mark.gpu

JS_RTOL = 1e-5           # north-star bar
JS_RTOL_EXPECT = 2e-7    # what the fp32 path is expected to deliver on these inputs


@pytest.fixture(scope="module")
def pk():
    import poppunk_mod_b200 as pkg

    assert pkg._native.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return pkg


@pytest.fixture(scope="module")
def ko():
    from oracle import kde_oracle

    return kde_oracle


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def grids(pk, a, b, kw, R, algo, h=0.03):
    return pk.KDE_distance._discrepancy(a, b, kw["gamma"], kw["eps"], kw["log"], R, h, "epanechnikov", algo, None,
                                        want_z=True, want_density=True)


@pytest.mark.parametrize("name", ALL_CASES)
def test_golden_cases_fp32(pk, golden, name):
    case = golden["cases"][name]
    a, b = golden_inputs(name)
    js = pk.KDE_JS_divergence(a, b, resolution=case["R"], algo="binned32", **case_kwargs(case))
    assert isinstance(js, np.float64)
    for key in ("oracle", "ref_masked"):
        want = case[key]
        assert abs(js - want) <= JS_RTOL * abs(want) + 1e-9, (name, key, js, want)
        assert abs(js - want) <= JS_RTOL_EXPECT * abs(want) + 1e-9, (name, key, js, want)


@pytest.mark.parametrize("name", ["G1_g025", "G5_ragged", "G1_R256", "G2_disjoint", "G3_identical", "G4_R100_g025",
                                  "G4_R256_g025", "G6_const_col", "G7_tiny"])
def test_per_node_density_fp32(pk, ko, golden, name):
    case = golden["cases"][name]
    a, b = golden_inputs(name)
    kw = case_kwargs(case)
    js, z1, z2, d1, d2 = grids(pk, a, b, kw, case["R"], "binned32")
    js64, y1, y2, e1, e2 = grids(pk, a, b, kw, case["R"], "binned")
    _, oz1, oz2, od1, od2 = ko.discrepancy(a, b, resolution=case["R"], return_grids=True, **kw)
    for got, want, exact in ((d1, od1, e1), (d2, od2, e2)):
        assert np.all(np.abs(got - want) <= 1e-6 * want + 1e-7 * want.max())
        assert np.array_equal(got == 0, want == 0)            # exact zeros stay exact zeros
        assert np.array_equal(got == 0, exact == 0)
    for got, want in ((z1, oz1), (z2, oz2)):
        # no per-node bound on the probabilities: **gamma amplifies the relative error of the tiniest densities
        # (a fringe node at 1e-7 of the maximum, off by 1e-7 of the maximum, moves its z**0.25 by 2e-3 relative);
        # what the path owes is the density bound above and the JS value below
        assert got.sum() == pytest.approx(1.0, abs=1e-12)
        assert np.array_equal(got == 0, want == 0)
    assert rel(js, js64) < JS_RTOL_EXPECT or abs(js - js64) < 1e-9


def test_batch_equals_singles_bitwise_fp32(pk, ko):
    from oracle import workloads as wl

    obs = wl.synthetic_target(20000)
    sims = [wl.sim_batch_member(s, n=5000 + 997 * s) for s in range(37)]
    got = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0, algo="binned32")
    exact = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0, algo="binned")
    np.testing.assert_allclose(got, exact, rtol=JS_RTOL_EXPECT)
    for i in (0, 5, 31, 32, 36):
        assert got[i] == pk.KDE_JS_divergence(obs, sims[i], gamma=0.25, eps=0.0, algo="binned32")
    np.testing.assert_array_equal(pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0, algo="binned32"), got)


@pytest.mark.parametrize("algo", ["binned32", "binned"])
def test_batch_equals_singles_bitwise_in_mixed_batches(pk, algo):
    """An evaluation gets the same arithmetic whatever else travels in its wave: sets too small for the fp32
    kernel (< 64 points: fp64 whatever the algorithm says) and, on a fine grid, sets dense enough for moment cells
    sit next to ordinary ones in one call (tools/stress.py found the first case with the default algorithm)."""
    from oracle import workloads as wl

    rng = np.random.default_rng(11)
    obs = wl.synthetic_target(9000)
    sims = [wl.sim_batch_member(s, n=int(n)) for s, n in enumerate([4000, 3, 700, 63, 64, 12000, 1, 2500, 40, 5000])]
    got = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0, resolution=100, algo=algo)
    for i in range(len(sims)):
        assert got[i] == pk.KDE_JS_divergence(obs, sims[i], gamma=0.25, eps=0.0, resolution=100, algo=algo), i
    # fine grid (128 x 128 cells): 128 points per cell on average = 2.1 M points per set is where moment cells start;
    # one evaluation above that bar between two below it
    big = wl.synthetic_target(2_200_000, seed=5)
    pair = [wl.sim_batch_member(1, n=30_000), wl.sim_S(7, 1800, 0.36, n=2_150_000), wl.sim_batch_member(2, n=50_000)]
    got = pk.KDE_JS_divergence_batch(big, pair, gamma=0.25, eps=0.0, resolution=640, algo=algo)
    for i in range(len(pair)):
        assert got[i] == pk.KDE_JS_divergence(big, pair[i], gamma=0.25, eps=0.0, resolution=640, algo=algo), i


def test_full_size_fp32(pk, ko):
    """BASELINE config 2 size on the reference's example target: oracle, identity, symmetry, disjoint."""
    from oracle import workloads as wl

    T = wl.example_target()
    S = wl.sim_S(254, 2500, 0.38)
    kw = dict(gamma=0.25, eps=0.0, resolution=256, algo="binned32")
    js = pk.KDE_JS_divergence(T, S, **kw)
    assert rel(js, ko.discrepancy(T, S, 0.25, 0.0, resolution=256)) < JS_RTOL_EXPECT
    assert pk.KDE_JS_divergence(T, T.copy(), **kw) == 0.0
    assert pk.KDE_JS_divergence(S, T, **kw) == pytest.approx(js, rel=1e-7)
    far = S + [10.0, 10.0]
    assert pk.KDE_JS_divergence(T, far, **kw) == pytest.approx(ko.SQRT_LN2, rel=1e-12)
    # the grid-search call site: gamma = 1e-12 turns every non-zero node into ~1, so a single node whose zero-ness
    # differed from the exact sum would move the result by ~1e-5: exact zeros are what keeps this one tight
    g = pk.KDE_JS_divergence(T, S, 1e-12, 1e-12, True, resolution=256, algo="binned32")
    assert rel(g, pk.KDE_JS_divergence(T, S, 1e-12, 1e-12, True, resolution=256, algo="binned")) < 1e-9


@pytest.mark.parametrize("R", [2, 3, 5, 17, 50, 99, 100, 104, 129, 256, 511, 1024])
def test_resolutions_fp32(pk, ko, R):
    from oracle import workloads as wl

    a, b = wl.ragged_inputs()
    got = pk.KDE_JS_divergence(a, b, gamma=0.5, eps=1e-12, resolution=R, algo="binned32")
    want = ko.discrepancy(a, b, 0.5, 1e-12, resolution=R)
    assert (np.isnan(got) and np.isnan(want)) or rel(got, want) < 1e-6


@pytest.mark.parametrize("h", [0.005, 0.02, 0.03, 0.2, 0.7])      # 0.005: outside the fixed-point format -> fp64 kernel
def test_bandwidths_fp32(pk, ko, h):
    from oracle import workloads as wl

    a, b = wl.g1_inputs()
    got = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, bandwidth=h, algo="binned32")
    assert rel(got, ko.discrepancy(a, b, 0.25, 0.0, bandwidth=h)) < 1e-6


def test_borderline_candidates_are_redone_in_fp64(pk, ko):
    """Points placed (to fp64 rounding) exactly on the edge of a node's support: the fp32 arithmetic cannot decide
    them, the tile goes to the fp64 redo list and the zero pattern must be the fp64 path's, node for node."""
    R, h = 41, 0.03                      # nodes at k/40
    rng = np.random.default_rng(5)
    nodes = rng.integers(5, 35, (400, 2)) / 40.0
    ang = rng.uniform(0, 2 * np.pi, 400)
    on_edge = nodes + h * np.column_stack([np.cos(ang), np.sin(ang)])            # distance h up to rounding
    axis = nodes + np.column_stack([np.full(400, h), np.zeros(400)])             # exactly representable offsets
    pts = np.vstack([on_edge, axis, [[0.0, 0.0], [1.0, 1.0]]])
    for algo_pts in (pts, pts[::-1].copy()):
        d32 = pk.kde_density(algo_pts, resolution=R, bandwidth=h, algo="binned32")
        d64 = pk.kde_density(algo_pts, resolution=R, bandwidth=h, algo="binned")
        assert np.array_equal(d32 == 0, d64 == 0)
        assert np.all(np.abs(d32 - d64) <= 1e-6 * d64 + 1e-7 * d64.max())
    want = ko.kde_density(pts, R, h).reshape(R, R)
    assert np.all(np.abs(d32 - want) <= 1e-6 * want + 1e-7 * want.max())


def test_randomised_configurations_fp32(pk, ko):
    rng = np.random.default_rng(20261020)
    worst = 0.0
    for trial in range(40):
        n1, n2 = int(rng.integers(1, 6000)), int(rng.integers(1, 6000))
        R = int(rng.choice([7, 20, 64, 100, 101, 130, 200]))
        h = float(rng.choice([0.02, 0.03, 0.03, 0.08, 0.3]))
        gamma = float(rng.choice([0.25, 0.25, 0.5, 1.0, 0.1]))
        eps = float(rng.choice([0.0, 1e-12]))
        spread = rng.choice([0.002, 0.05, 1.0])
        a = rng.normal([0.5, 0.5], spread, (n1, 2))
        b = rng.normal([0.5 + spread, 0.5], spread * rng.uniform(0.5, 2), (n2, 2))
        if trial % 5 == 0:
            a = np.round(a, 3)                                # heavy duplication (Pansim-like discreteness)
        got = pk.KDE_JS_divergence(a, b, gamma, eps, resolution=R, bandwidth=h, algo="binned32")
        exact = pk.KDE_JS_divergence(a, b, gamma, eps, resolution=R, bandwidth=h, algo="binned")
        assert abs(got - exact) <= 1e-6 * exact + 1e-9, (trial, got, exact)
        worst = max(worst, abs(got - exact) / max(exact, 1e-12))
    print(f"worst relative gap fp32 vs fp64 culled path: {worst:.2e}")


def test_fine_grid_moments_fp32(pk, ko, monkeypatch):
    """R = 1024 with dense cells: moment cells (fp64) + fp32 point runs."""
    from oracle import workloads as wl

    a, b = wl.synthetic_target(300_000, seed=11), wl.sim_S(5, 1800, 0.36, n=300_000)
    monkeypatch.setenv("KDEJS_MOMENTS", "1")
    exact = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=1024, algo="binned")
    # by default moment geometries stay on the fp64 kernel whatever the algorithm says ...
    assert pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=1024, algo="binned32") == exact
    # ... KDEJS_FP32=1 forces the fp32 kernel (moment cells in fp64, point runs in fp32)
    monkeypatch.setenv("KDEJS_FP32", "1")
    got = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=1024, algo="binned32")
    assert got != exact and rel(got, exact) < JS_RTOL_EXPECT


def test_device_entry_and_profile_slots_fp32(pk):
    import torch

    from oracle import workloads as wl

    nat = pk._native
    obs = wl.synthetic_target(20000)
    sims = [wl.sim_batch_member(s, n=15000) for s in range(5)]
    want = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0, resolution=128, algo="binned32")
    dobs = torch.from_numpy(obs).cuda()
    dsims = [torch.from_numpy(s).cuda() for s in sims]
    out = torch.empty(5, dtype=torch.float64, device="cuda")
    ptrs = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in dsims])
    cnt = (ctypes.c_int64 * 5)(*[t.shape[0] for t in dsims])
    L = nat.lib()
    nat.check(L.kdejs_profile_reset(0))
    nat.check(L.kdejs_profile_enable(0, 1))
    nat.check(L.kdejs_discrepancy_batch_dev(0, dobs.data_ptr(), len(obs), ptrs, cnt, 5, 128, 0.03, 0.25, 0.0, 0, 0,
                                            nat.ALGOS["binned32"], out.data_ptr(), None,
                                            torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    ms, n = ctypes.c_double(), ctypes.c_int64()
    nat.check(L.kdejs_profile_read(0, nat.KERNEL_NAMES.index("redo64"), ctypes.byref(ms), ctypes.byref(n)))
    assert n.value == 1
    nat.check(L.kdejs_profile_enable(0, 0))
