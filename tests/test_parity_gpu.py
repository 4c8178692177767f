"""GPU: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors."""
import numpy as np
import pytest

from conftest import case_kwargs, golden_inputs

pytestmark = pytest.mark.gpu

ALL_CASES = ["G1_default", "G1_g025", "G1_gridsearch", "G1_nan", "G1_R256", "G1_R37", "G2_disjoint",
             "G2_disjoint_default", "G3_identical", "G5_ragged", "G5_ragged_default", "G6_const_col", "G7_tiny",
             "G4_R100_g025", "G4_R100_g1", "G4_R256_g025", "G4_gridsearch", "C3_member0", "C3_member1",
             "C3_member2", "C3_member3"]
JS_RTOL = 1e-5          # the north-star bar (BASELINE.json); observed ~1e-13
JS_RTOL_EXPECT = 1e-10  # what the fp64 path actually has to deliver


@pytest.fixture(scope="module")
def pk():
    """This module pins the EXACT fp64 culled sum (algo="binned") wherever a test does not name an algorithm: its
    bars are 1e-10 and bitwise equalities.  The package default (the fp32 culled kernel) is covered, with fp32
    bars, by tests/test_binned32_gpu.py."""
    import poppunk_mod_b200 as pkg

    assert pkg._native.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    old = pkg.KDE_distance.DEFAULT_ALGO
    pkg.KDE_distance.DEFAULT_ALGO = "binned"
    yield pkg
    pkg.KDE_distance.DEFAULT_ALGO = old


@pytest.fixture(scope="module")
def ko():
    from oracle import kde_oracle

    return kde_oracle


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


@pytest.mark.parametrize("name", ALL_CASES)
def test_golden_cases(pk, ko, golden, name):
    case = golden["cases"][name]
    a, b = golden_inputs(name)
    a0, b0 = a.copy(), b.copy()
    js = pk.KDE_JS_divergence(a, b, resolution=case["R"], **case_kwargs(case))
    assert isinstance(js, np.float64)
    np.testing.assert_array_equal(a, a0)
    np.testing.assert_array_equal(b, b0)
    for key in ("oracle", "ref_masked"):
        want = case[key]
        assert abs(js - want) <= JS_RTOL_EXPECT * abs(want) + 1e-12, (name, key, js, want)
        assert abs(js - want) <= JS_RTOL * abs(want) + 1e-12
    if case["gamma"] == 1.0:      # raw reference is residue-free here
        assert abs(js - case["ref_raw"]) <= 1e-9 * abs(case["ref_raw"]) + 1e-12


@pytest.mark.parametrize("name", ["G1_g025", "G5_ragged", "G4_R100_g025"])
def test_per_node_density_and_probabilities(pk, ko, golden, golden_grids, name):
    case = golden["cases"][name]
    a, b = golden_inputs(name)
    kw = case_kwargs(case)
    js, z1, z2, d1, d2 = pk.KDE_distance._discrepancy(a, b, kw["gamma"], kw["eps"], kw["log"], case["R"], 0.03,
                                                      "epanechnikov", None, None, want_z=True, want_density=True)
    _, oz1, oz2, od1, od2 = ko.discrepancy(a, b, resolution=case["R"], return_grids=True, **kw)
    for got, want in ((d1, od1), (d2, od2)):
        # SURVEY.md 8c(i), fp64 kernel: |dz| <= 1e-6 z + 1e-12 z_max   (observed ~1e-15 z_max)
        assert np.all(np.abs(got - want) <= 1e-6 * want + 1e-12 * want.max())
        assert np.array_equal(got == 0, want == 0)            # exact zeros stay exact zeros
    for got, want in ((z1, oz1), (z2, oz2)):
        assert np.all(np.abs(got - want) <= 1e-6 * want + 1e-12 * want.max())
        assert got.sum() == pytest.approx(1.0, abs=1e-12)
    for got, key in ((d1, "dens1"), (d2, "dens2")):           # 8c(ii): masked reference
        ref = golden_grids[f"{name}_{key}"]
        nz = ref > 0
        assert np.all(np.abs(got[nz] - ref[nz]) <= 1e-6 * ref[nz] + 1e-11 * ref.max())
    z1b, z2b, grid = pk.scale_KDE(a, b, kw["gamma"], kw["eps"], log=kw["log"], resolution=case["R"])
    np.testing.assert_array_equal(z1b, z1)
    assert grid[2].shape == (case["R"] ** 2, 2)


def test_batch_equals_singles_bitwise(pk, ko):
    from oracle import workloads as wl

    obs = wl.synthetic_target(20000)
    sims = [wl.sim_batch_member(s, n=5000 + 997 * s) for s in range(37)]       # ragged sizes, > one wave
    got = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0)
    assert got.shape == (37,) and got.dtype == np.float64
    for i in (0, 5, 31, 32, 36):
        single = pk.KDE_JS_divergence(obs, sims[i], gamma=0.25, eps=0.0)
        assert got[i] == single                                                # same kernels, same order
        want = ko.discrepancy(obs, sims[i], 0.25, 0.0)
        assert rel(got[i], want) < JS_RTOL_EXPECT
    again = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0)
    np.testing.assert_array_equal(again, got)                                  # run-to-run reproducible


def test_batch_accepts_one_3d_array(pk):
    """sims as ONE (B,N,2) array (an ELFI batch of equal-sized simulations): same values as the list of its rows, for
    contiguous, strided and non-float64 inputs, on one device and dealt over the visible ones; summary_batch too."""
    from oracle import workloads as wl

    obs = wl.synthetic_target(8000)
    arena = np.stack([wl.sim_batch_member(s, n=6000) for s in range(21)])
    want = pk.KDE_JS_divergence_batch(obs, [arena[i] for i in range(21)], gamma=0.25, eps=0.0)
    np.testing.assert_array_equal(pk.KDE_JS_divergence_batch(obs, arena, gamma=0.25, eps=0.0), want)
    np.testing.assert_array_equal(pk.KDE_JS_divergence_batch(obs, arena[::2], gamma=0.25, eps=0.0), want[::2])
    np.testing.assert_array_equal(pk.KDE_JS_divergence_batch(obs, arena[3:11], gamma=0.25, eps=0.0), want[3:11])
    wide = np.zeros((21, 6000, 3)); wide[:, :, :2] = arena                      # rows not 16 bytes apart: copied
    np.testing.assert_array_equal(pk.KDE_JS_divergence_batch(obs, wide[:, :, :2], gamma=0.25, eps=0.0), want)
    assert pk.KDE_JS_divergence_batch(obs, arena[:0], gamma=0.25, eps=0.0).shape == (0,)
    devs = list(range(pk._native.device_count()))
    np.testing.assert_array_equal(pk.KDE_JS_divergence_batch(obs, arena, gamma=0.25, eps=0.0, devices=devs), want)
    freqs = [np.linspace(0.0, 1.0, 50 + i) for i in range(21)]
    observed = [0.0, 0.3, 0.3, 0.4]
    s_list = pk.KDE_distance.summary_batch(obs, [arena[i] for i in range(21)], freqs, observed, 0.25, 0.95, 0.15)
    s_3d = pk.KDE_distance.summary_batch(obs, arena, freqs, observed, 0.25, 0.95, 0.15)
    for a, b in zip(s_list, s_3d):
        np.testing.assert_array_equal(a, b)
    with pytest.raises(ValueError):
        pk.KDE_JS_divergence_batch(obs, np.zeros((4, 10, 3)), gamma=0.25, eps=0.0)


def test_pairs_entry(pk, ko):
    from oracle import workloads as wl

    sets = [wl.sim_batch_member(s, n=3000 + 500 * s) for s in range(5)]
    pairs = [(i, j) for i in range(5) for j in range(i + 1, 5)]
    got = pk.KDE_JS_divergence_pairs(sets, pairs, gamma=0.25, eps=0.0)
    for k, (i, j) in enumerate(pairs):
        assert rel(got[k], ko.discrepancy(sets[i], sets[j], 0.25, 0.0)) < JS_RTOL_EXPECT
    sym = pk.KDE_JS_divergence_pairs(sets, [(j, i) for i, j in pairs], gamma=0.25, eps=0.0)
    np.testing.assert_allclose(sym, got, rtol=1e-13)


def test_full_size_properties(pk, ko):
    """BASELINE config 2 size (100k vs 100k, R=256): oracle check plus size-independent properties."""
    from oracle import workloads as wl

    T = wl.synthetic_target(100_000)
    S = wl.sim_S(254, 2500, 0.38)
    js = pk.KDE_JS_divergence(T, S, gamma=0.25, eps=0.0, resolution=256)
    assert rel(js, ko.discrepancy(T, S, 0.25, 0.0, resolution=256)) < JS_RTOL_EXPECT
    assert 0.0 < js < ko.SQRT_LN2
    assert pk.KDE_JS_divergence(T, T.copy(), gamma=0.25, eps=0.0, resolution=256) == 0.0     # identity
    assert pk.KDE_JS_divergence(S, T, gamma=0.25, eps=0.0, resolution=256) == pytest.approx(js, rel=1e-13)
    perm = np.random.default_rng(0).permutation(len(S))
    assert pk.KDE_JS_divergence(T, S[perm], gamma=0.25, eps=0.0, resolution=256) == pytest.approx(js, rel=1e-12)
    far = S + [10.0, 10.0]                                                     # disjoint after joint scaling
    assert pk.KDE_JS_divergence(T, far, gamma=0.25, eps=0.0, resolution=256) == pytest.approx(ko.SQRT_LN2, rel=1e-14)
    z1, z2, _ = pk.scale_KDE(T, S, 0.25, 0.0, resolution=256)
    assert z1.sum() == pytest.approx(1.0, abs=1e-12) and z1.min() >= 0.0
    # affine invariance of the joint min-max scaling
    js_aff = pk.KDE_JS_divergence(T * [2.0, 0.5] + [1.0, -3.0], S * [2.0, 0.5] + [1.0, -3.0], gamma=0.25, eps=0.0,
                                  resolution=256)
    assert js_aff == pytest.approx(js, rel=1e-9)


# 50, 99, 104: resolutions where (R-1) * fl(1/(R-1)) != 1.0 while numpy.linspace forces the last node to the maximum
@pytest.mark.parametrize("R", [1, 2, 3, 5, 17, 50, 99, 100, 104, 129, 256, 511, 1024, 1500])     # >= 1024: 128x128 culling cells
def test_resolutions(pk, ko, R):
    from oracle import workloads as wl

    a, b = wl.ragged_inputs()
    got = pk.KDE_JS_divergence(a, b, gamma=0.5, eps=1e-12, resolution=R)
    want = ko.discrepancy(a, b, 0.5, 1e-12, resolution=R)
    assert (np.isnan(got) and np.isnan(want)) or rel(got, want) < JS_RTOL_EXPECT


@pytest.mark.parametrize("h", [0.005, 0.03, 0.2, 0.7])
def test_bandwidths(pk, ko, h):
    from oracle import workloads as wl

    a, b = wl.g1_inputs()
    got = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, bandwidth=h)
    assert rel(got, ko.discrepancy(a, b, 0.25, 0.0, bandwidth=h)) < JS_RTOL_EXPECT


def test_dense64_matches_binned(pk, ko):
    from oracle import workloads as wl

    a, b = wl.ragged_inputs()
    j0 = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, algo="binned")
    j1 = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, algo="dense64")
    assert rel(j1, j0) < 1e-11


def test_gaussian_kernel(pk, ko):
    from oracle import workloads as wl

    a, b = wl.g1_inputs()
    got = pk.KDE_JS_divergence(a[:800], b[:900], gamma=1.0, eps=1e-12, kernel="gaussian", resolution=64)
    want = ko.discrepancy(a[:800], b[:900], 1.0, 1e-12, kernel="gaussian", resolution=64)
    assert rel(got, want) < 1e-9


def test_kde_density(pk, ko):
    rng = np.random.default_rng(8)
    X = rng.random((5000, 2))
    got = pk.kde_density(X, resolution=100)
    want = ko.kde_density(X, 100, 0.03).reshape(100, 100)
    assert np.all(np.abs(got - want) <= 1e-6 * want + 1e-12 * want.max())
    # off-unit grid and points outside it (plot_scatter scales by the maximum only)
    Y = rng.normal(0.5, 0.4, (3000, 2))
    got = pk.kde_density(Y, resolution=80, bandwidth=0.05, minimum=-0.2, maximum=1.3)
    want = ko.kde_density(Y, 80, 0.05, gmin=-0.2, gmax=1.3).reshape(80, 80)
    assert np.all(np.abs(got - want) <= 1e-6 * want + 1e-12 * want.max())
    with pytest.raises(ValueError):
        pk.kde_density(np.array([[np.nan, 0.1]]))


def test_nan_and_infinity_handling(pk, ko):
    a, b = golden_inputs("G1_default")
    a = a.copy()
    a[::7, 0] = np.nan
    a[3] = np.nan
    got = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0)
    assert rel(got, ko.discrepancy(a, b, 0.25, 0.0)) < JS_RTOL_EXPECT
    b2 = b.copy()
    b2[10, 1] = np.inf
    with pytest.raises(ValueError, match="infinity"):
        pk.KDE_JS_divergence(a, b2)
    with pytest.raises(ValueError, match="infinity"):
        pk.KDE_JS_divergence(np.zeros((4, 2)), b, log=True, eps=0.0)
    out = None
    try:
        pk.KDE_JS_divergence_batch(a, [b, b2, b], gamma=0.25, eps=0.0)
    except ValueError as e:
        out = str(e)
    assert out and "infinity" in out
    # all-NaN column: every scaled value becomes 0 (sklearn: NaN scale, then KDE_distance.py:53)
    c = b.copy()
    c[:, 0] = np.nan
    a3 = a.copy()
    a3[:, 0] = np.nan
    got = pk.KDE_JS_divergence(a3, c, gamma=1.0, eps=1e-12)
    assert rel(got, ko.discrepancy(a3, c, 1.0, 1e-12)) < JS_RTOL_EXPECT


def test_log_mode_and_kl(pk, ko):
    a, b = golden_inputs("G1_default")
    got = pk.KDE_JS_divergence(a, b, 1e-12, 1e-12, True)            # positional, as distance_to_gridsearch.py:50
    assert rel(got, ko.discrepancy(a, b, 1e-12, 1e-12, True)) < JS_RTOL_EXPECT
    kl = pk.KDE_KL_divergence(a, b)
    _, z1, z2, _, _ = ko.discrepancy(a, b, 1.0, 1e-12, return_grids=True)
    assert kl == pytest.approx(float(np.sum(z1 * np.log(z1 / z2))), rel=1e-9)


def test_row_sharded_phases_reproduce_the_full_result(pk, ko):
    import ctypes

    from oracle import workloads as wl

    nat = pk._native
    a, b = wl.synthetic_target(30000), wl.sim_S(3, 1500, 0.33, n=25000)
    R = 130
    full = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=R)
    bounds = [0, 32, 64, 100, R]
    norms, lrs = [], []
    common = (R, 0.03, 0.25, 0.0, 0, 0, 0)
    for lo, hi in zip(bounds[:-1], bounds[1:]):     # one "rank" at a time on the single test GPU
        n = (ctypes.c_double * 2)()
        nat.check(nat.lib().kdejs_shard_phase1(0, nat.ptr(a), len(a), nat.ptr(b), len(b), *common, lo, hi, n))
        norms.append((n[0], n[1]))
    tot = (ctypes.c_double * 2)(sum(x for x, _ in norms), sum(y for _, y in norms))
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        n = (ctypes.c_double * 2)()
        lr = (ctypes.c_double * 2)()
        nat.check(nat.lib().kdejs_shard_phase1(0, nat.ptr(a), len(a), nat.ptr(b), len(b), *common, lo, hi, n))
        nat.check(nat.lib().kdejs_shard_phase2(0, tot, lr))
        lrs.append((lr[0], lr[1]))
    js = np.sqrt((sum(x for x, _ in lrs) + sum(y for _, y in lrs)) / 2.0)
    assert js == pytest.approx(full, rel=1e-12)
    assert rel(js, ko.discrepancy(a, b, 0.25, 0.0, resolution=R)) < JS_RTOL_EXPECT


def test_device_resident_entry_matches_host_entry(pk):
    import ctypes

    import torch

    from oracle import workloads as wl

    nat = pk._native
    obs = wl.synthetic_target(20000)
    sims = [wl.sim_batch_member(s, n=15000) for s in range(5)]
    want = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0, resolution=128)
    dobs = torch.from_numpy(obs).cuda()
    dsims = [torch.from_numpy(s).cuda() for s in sims]
    out = torch.empty(5, dtype=torch.float64, device="cuda")
    status = torch.empty(5, dtype=torch.int32, device="cuda")
    ptrs = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in dsims])
    cnt = (ctypes.c_int64 * 5)(*[t.shape[0] for t in dsims])
    st = torch.cuda.current_stream().cuda_stream
    nat.check(nat.lib().kdejs_discrepancy_batch_dev(0, dobs.data_ptr(), len(obs), ptrs, cnt, 5, 128, 0.03, 0.25,
                                                    0.0, 0, 0, 0, out.data_ptr(), status.data_ptr(), st))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    assert status.cpu().numpy().tolist() == [0] * 5


def test_profile_counters(pk):
    import ctypes

    nat = pk._native
    a, b = golden_inputs("G1_default")
    L = nat.lib()
    nat.check(L.kdejs_profile_reset(0))
    nat.check(L.kdejs_profile_enable(0, 1))
    n0 = ctypes.c_int64()
    nat.check(L.kdejs_launch_count(0, ctypes.byref(n0), 1))
    pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0)
    ms, n = ctypes.c_double(), ctypes.c_int64()
    nat.check(L.kdejs_profile_read(0, 4, ctypes.byref(ms), ctypes.byref(n)))
    assert n.value == 1 and ms.value > 0.0
    cnt = ctypes.c_int64()
    nat.check(L.kdejs_launch_count(0, ctypes.byref(cnt), 0))
    assert cnt.value == 8
    w = ctypes.c_double()
    nat.check(L.kdejs_work_counters(0, ctypes.byref(w), 1))
    assert w.value > 0
    nat.check(L.kdejs_profile_enable(0, 0))


# ---- the dense fp32 kernel (north-star brute force): fp32 tolerance of SURVEY.md 8c(i) -------------
@pytest.mark.parametrize("name", ["G1_g025", "G5_ragged", "G1_R256", "G2_disjoint", "G3_identical", "G4_R100_g025"])
def test_dense_fp32_kernel(pk, ko, golden, name):
    case = golden["cases"][name]
    a, b = golden_inputs(name)
    kw = case_kwargs(case)
    js, z1, z2, d1, d2 = pk.KDE_distance._discrepancy(a, b, kw["gamma"], kw["eps"], kw["log"], case["R"], 0.03,
                                                      "epanechnikov", "dense", None, want_z=True, want_density=True)
    _, oz1, oz2, od1, od2 = ko.discrepancy(a, b, resolution=case["R"], return_grids=True, **kw)
    assert abs(js - case["oracle"]) <= 1e-5 * case["oracle"] + 1e-7          # the north-star bar for the JS value
    for got, want in ((d1, od1), (d2, od2)):
        assert np.all(np.abs(got - want) <= 1e-6 * want + 1e-7 * want.max())  # fp32 kernel: absolute floor 1e-7 z_max


def test_dense_fp32_gaussian_and_batch(pk, ko):
    from oracle import workloads as wl

    a, b = wl.g1_inputs()
    got = pk.KDE_JS_divergence(a[:800], b[:900], gamma=1.0, eps=1e-12, kernel="gaussian", resolution=64, algo="dense")
    want = ko.discrepancy(a[:800], b[:900], 1.0, 1e-12, kernel="gaussian", resolution=64)
    assert rel(got, want) < 1e-5
    obs = wl.synthetic_target(20000)
    sims = [wl.sim_batch_member(s, n=3000 + 700 * s) for s in range(5)]
    dense = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0, algo="dense")
    exact = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0)
    np.testing.assert_allclose(dense, exact, rtol=1e-5)


def test_randomised_configurations(pk, ko):
    """40 seeded random configurations (sizes, resolution, bandwidth, gamma, eps, log, NaNs, duplicates,
    clustered vs spread clouds) against the exact oracle."""
    rng = np.random.default_rng(20261019)
    worst = 0.0
    for trial in range(40):
        n1, n2 = int(rng.integers(1, 6000)), int(rng.integers(1, 6000))
        R = int(rng.choice([7, 20, 64, 100, 101, 130, 200]))
        h = float(rng.choice([0.01, 0.03, 0.03, 0.08, 0.3]))
        gamma = float(rng.choice([0.25, 0.25, 0.5, 1.0, 0.1, 2.0]))
        eps = float(rng.choice([0.0, 0.0, 1e-12, 1e-6]))
        log = bool(rng.integers(0, 4) == 0)
        spread = rng.choice([1e-4, 1e-2, 0.3])
        a = np.abs(rng.normal([0.002, 0.3], [spread * 0.01, spread], (n1, 2))) + 1e-9
        b = np.abs(rng.normal([0.0025, 0.28], [spread * 0.02, spread * 1.5], (n2, 2))) + 1e-9
        if rng.integers(0, 3) == 0:
            a[rng.integers(0, n1, max(1, n1 // 20))] = a[0]            # duplicates
        if rng.integers(0, 3) == 0:
            b[rng.integers(0, n2, max(1, n2 // 10)), rng.integers(0, 2)] = np.nan
        if rng.integers(0, 5) == 0:
            b[:, 0] = np.round(b[:, 0], 4)
        got = pk.KDE_JS_divergence(a, b, gamma, eps, log, resolution=R, bandwidth=h)
        want = ko.discrepancy(a, b, gamma, eps, log, resolution=R, bandwidth=h)
        if np.isnan(want):
            assert np.isnan(got), (trial, got, want)
            continue
        err = abs(got - want) / max(abs(want), 1e-12)
        worst = max(worst, err)
        assert err < 1e-9, (trial, n1, n2, R, h, gamma, eps, log, got, want)
    print("worst relative error over random configurations:", worst)


def test_work_balanced_row_bands(pk, ko):
    """kdejs_shard_phase1_balanced: bands are contiguous, cover all rows, and reproduce the full result."""
    import ctypes

    from oracle import workloads as wl

    nat = pk._native
    a, b = wl.synthetic_target(40000), wl.sim_S(9, 2200, 0.3, n=30000)
    R, world = 200, 3
    common = (R, 0.03, 0.25, 0.0, 0, 0, 0)
    full = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=R)

    def phase1(rank):
        n, rows = (ctypes.c_double * 2)(), (ctypes.c_int * 2)()
        nat.check(nat.lib().kdejs_shard_phase1_balanced(0, nat.ptr(a), len(a), nat.ptr(b), len(b), 0, *common,
                                                        rank, world, n, rows))
        return (n[0], n[1]), (rows[0], rows[1])

    norms, bands = zip(*[phase1(r) for r in range(world)])
    assert bands[0][0] == 0 and bands[-1][1] == R
    for (lo, hi), (lo2, _) in zip(bands[:-1], bands[1:]):
        assert hi == lo2 and lo <= hi
    sizes = [hi - lo for lo, hi in bands]
    assert max(sizes) > min(sizes)                 # concentrated data: equal work is not equal row counts
    tot = (ctypes.c_double * 2)(sum(x for x, _ in norms), sum(y for _, y in norms))
    L = Rr = 0.0
    for r in range(world):
        phase1(r)
        lr = (ctypes.c_double * 2)()
        nat.check(nat.lib().kdejs_shard_phase2(0, tot, lr))
        L += lr[0]
        Rr += lr[1]
    js = np.sqrt((L + Rr) / 2.0)
    assert js == pytest.approx(full, rel=1e-12)
    assert rel(js, ko.discrepancy(a, b, 0.25, 0.0, resolution=R)) < JS_RTOL_EXPECT


def _fork_child(q):
    import numpy as _np

    import poppunk_mod_b200 as _pk
    from oracle import workloads as _wl

    a, b = _wl.g1_inputs()
    q.put(float(_pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0)))


def test_fork_after_import_before_first_use(golden):
    """ELFI's multiprocessing client and the scripts' Pools fork AFTER importing KDE_distance
    (poppunk_mod.py:427-429, pairwise_2D_distance.py:116): loading the library must not touch CUDA, so a
    forked worker can create its own context.  Runs in a fresh interpreter so that no CUDA context exists
    in the forking process."""
    import subprocess
    import sys
    import textwrap

    from conftest import ROOT

    code = textwrap.dedent(f"""
        import sys, multiprocessing as mp
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {ROOT + '/tests'!r})
        import poppunk_mod_b200 as pk            # dlopen only
        pk._native.lib()
        from test_parity_gpu import _fork_child
        ctx = mp.get_context("fork")
        q = ctx.Queue()
        ps = [ctx.Process(target=_fork_child, args=(q,)) for _ in range(2)]
        [p.start() for p in ps]
        vals = [q.get(timeout=120) for _ in ps]
        [p.join(60) for p in ps]
        assert all(p.exitcode == 0 for p in ps), [p.exitcode for p in ps]
        print(vals[0], vals[1])
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    v0, v1 = (float(x) for x in out.stdout.split()[-2:])
    want = golden["cases"]["G1_g025"]["oracle"]
    assert v0 == v1 and abs(v0 - want) <= 1e-7 * want       # the child runs the package default (fp32 culled kernel)


def test_sweep_drivers_and_simulator_summary(pk, ko, tmp_path):
    """scripts/pairwise_2D_distance.py, scripts/distance_to_gridsearch.py and process_result on files."""
    import math

    from oracle import workloads as wl
    from poppunk_mod_b200 import simulator, sweeps

    sets = [wl.sim_batch_member(s, n=2000 + 300 * s) for s in range(4)]
    paths = []
    for i, a in enumerate(sets):
        p = tmp_path / f"species{i}.txt"
        if i % 2:      # 4-column PopPUNK table: two id columns first
            p.write_text("".join(f"A{k}\tB{k}\t{float(x)!r}\t{float(y)!r}\n" for k, (x, y) in enumerate(a)))
        else:
            np.savetxt(p, a, delimiter="\t", fmt="%.17g")
        paths.append(str(p))
    rows = sweeps.pairwise_distances(paths)
    assert [(r[0], r[1]) for r in rows] == [(paths[i], paths[j]) for i in range(4) for j in range(i + 1, 4)]
    k = 0
    for i in range(4):
        for j in range(i + 1, 4):
            assert rel(rows[k][2], ko.discrepancy(sets[i], sets[j], 0.25, 0.0)) < JS_RTOL_EXPECT
            assert rows[k][3] == (sets[j][:, 0].min(), sets[j][:, 0].max(), sets[j][:, 1].min(), sets[j][:, 1].max())
            k += 1
    with_fracs = [(paths[0], 0.3, 0.2, 0.5), (paths[1], 0.25, 0.25, 0.5)]
    d = sweeps.pairwise_distances(with_fracs)[0][2]
    js = ko.discrepancy(sets[0], sets[1], 0.25, 0.0)
    assert d == pytest.approx(math.dist((0, 0.3, 0.2, 0.5), (js, 0.25, 0.25, 0.5)), rel=1e-12)

    # grid search: reference = set 0, candidates = the others; the reference's call is (1e-12, log=True)
    best, dists = sweeps.gridsearch_best(paths[0], paths[1:], chunk=2)
    want = [ko.discrepancy(sets[0], s, 1e-12, 1e-12, True) for s in sets[1:]]
    np.testing.assert_allclose(dists, want, rtol=1e-9)
    assert best[0] == paths[1 + int(np.argmin(want))]
    none, _ = sweeps.gridsearch_best(paths[0], paths[1:], core_min_bounds=[9, 10], core_max_bounds=[9, 10])
    assert none == (None, 1.0)

    freqs = np.array([0.0, 0.02, 0.5, 0.97, 1.0])
    vec = simulator.process_result(paths[2], freqs, paths[0], 0.25, 0.95, 0.05)
    assert rel(vec[0], ko.discrepancy(sets[0][1:] if False else sets[0], sets[2], 0.25, 0.0)) < JS_RTOL_EXPECT
    assert tuple(vec[1:]) == (2 / 4, 1 / 4, 1 / 4)
    batch = simulator.process_results_batch([paths[2], sets[3]], [freqs, freqs], sets[0], 0.25, 0.95, 0.05)
    assert batch.shape == (2, 4) and batch[0, 0] == vec[0]


def test_summary_vector_and_elfi_nodes_on_device(pk, ko):
    """kdejs_summary_batch against a numpy restatement of poppunk_mod.py:285-291 (gene-frequency classes), :379
    (summary vector), :494 / :571-572 (euclidean distance to the observed vector, log)."""
    from oracle import workloads as wl
    from poppunk_mod_b200 import simulator

    rng = np.random.default_rng(12)
    obs = wl.synthetic_target(9000)
    sims = [wl.sim_batch_member(s, n=4000 + 500 * s) for s in range(5)]
    freqs = [np.round(rng.beta(0.3, 0.3, 3000 + 101 * s), 3) * (rng.random(3000 + 101 * s) > 0.2) for s in range(5)]
    freqs[3][:5] = [0.05, 0.95, 0.0, np.nan, 1.0]                       # values exactly on the thresholds, a NaN
    core, rare = 0.95, 0.05
    observed = np.array([0.0, 0.31, 0.22, 0.47])
    summary, d, log_d = pk.KDE_distance.summary_batch(obs, sims, freqs, observed, 0.25, core, rare)
    js = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0)
    for i, X in enumerate(freqs):
        with np.errstate(invalid="ignore"):
            genes = (X > 0).sum()
            want = np.array([js[i], (X >= core).sum() / genes, np.count_nonzero((rare <= X) & (X < core)) / genes,
                             np.count_nonzero((0 < X) & (X < rare)) / genes])
        np.testing.assert_array_equal(summary[i], want)                 # integer counts, IEEE divisions: exact
        dd = np.sqrt(((want - observed) ** 2).sum())
        assert d[i] == pytest.approx(dd, rel=1e-15) and log_d[i] == pytest.approx(np.log(dd), rel=1e-14)
    s2, d2, l2 = simulator.process_results_batch(sims, freqs, obs, 0.25, core, rare, observed=observed)
    np.testing.assert_array_equal(s2, summary)
    np.testing.assert_array_equal(l2, log_d)
    np.testing.assert_allclose(simulator.log_euclidean_discrepancy(summary, observed), log_d, rtol=1e-14)
    one = simulator.process_result(sims[2], freqs[2], obs, 0.25, core, rare)
    np.testing.assert_array_equal(one, summary[2])


def test_rule_of_thumb_bandwidths_against_sklearn(pk):
    """bandwidth="scott" | "silverman" (sklearn:neighbors/_kde.py:223-229) on the Gaussian kernel."""
    from sklearn.neighbors import KernelDensity

    rng = np.random.default_rng(3)
    X = rng.random((700, 2))
    R = 40
    _, _, xy = pk.get_grid(0, 1, R)
    for rule in ("scott", "silverman"):
        want = np.exp(KernelDensity(bandwidth=rule, kernel="gaussian").fit(X).score_samples(xy)).reshape(R, R)
        got = pk.kde_density(X, resolution=R, bandwidth=rule, kernel="gaussian")
        np.testing.assert_allclose(got, want, rtol=1e-9)
        assert pk.KDE_distance.rule_bandwidth(rule, 700) == pytest.approx(KernelDensity(bandwidth=rule).fit(X).bandwidth_)
    Y = rng.random((700, 2)) * 0.8
    a = pk.KDE_JS_divergence(X, Y, bandwidth="scott")                    # equal sizes: one bandwidth for both sets
    assert a == pk.KDE_JS_divergence(X, Y, bandwidth=700 ** (-1 / 6))
    with pytest.raises(ValueError, match="equal size"):
        pk.KDE_JS_divergence(X, Y[:500], bandwidth="silverman")
    with pytest.raises(ValueError):
        pk.kde_density(X, bandwidth="botev")


def test_large_evaluation_against_oracle(pk, ko):
    """1 M + 0.7 M pairs on a 512^2 grid (one wave of one, large sort chunks, many heavy tiles) and the
    chunk-size switch of the sort (kMaxChw) at 3 M points."""
    from oracle import workloads as wl

    a = wl.synthetic_target(1_000_000, seed=21)
    b = wl.sim_S(17, 1500, 0.33, n=700_000)
    got = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=512)
    want = ko.discrepancy(a, b, 0.25, 0.0, resolution=512)
    assert rel(got, want) < JS_RTOL_EXPECT
    big = wl.sim_S(3, 900, 0.4, n=3_000_000)
    got = pk.KDE_JS_divergence(big, b, gamma=1.0, eps=1e-12, resolution=100)
    want = ko.discrepancy(big, b, 1.0, 1e-12, resolution=100)
    assert rel(got, want) < JS_RTOL_EXPECT


def test_command_line_entry_point(pk, ko, tmp_path, capsys):
    """python KDE_distance.py --infile1 A --infile2 B prints `js_distance_no_eps: <value>` (KDE_distance.py:132-144)."""
    from oracle import workloads as wl

    a, b = wl.ragged_inputs()
    a = np.nan_to_num(a, nan=0.001)
    b = np.nan_to_num(b, nan=0.2)
    np.savetxt(tmp_path / "a.tsv", a, delimiter="\t", fmt="%.17g")
    np.savetxt(tmp_path / "b.tsv", b, delimiter="\t", fmt="%.17g")
    pk.KDE_distance.main(["--infile1", str(tmp_path / "a.tsv"), "--infile2", str(tmp_path / "b.tsv")])
    out = capsys.readouterr().out.strip()
    assert out.startswith("js_distance_no_eps: ")
    assert rel(float(out.split(": ")[1]), ko.discrepancy(a, b, 0.25, 0.0)) < JS_RTOL_EXPECT


def test_near_identical_distributions_and_wide_dynamic_range(pk, ko):
    """Two regressions found by tools/stress.py.  (1) Almost identical clouds: scipy's left/right sums cancel to
    ~1e-17 whose sign is noise (sqrt -> NaN or ~1e-9); the product sums rel_entr per node (>= 0) with log1p where
    p ~ m and must return a small finite number.  (2) gamma = 3 spreads z over > 16 decades: q/m < 1e-16 must still
    go through log(q/m)."""
    rng = np.random.default_rng(33)
    a = np.column_stack([rng.integers(1, 40, 29575) / 1e4, rng.integers(1, 60, 29575) / 100.0])
    b = np.column_stack([rng.integers(1, 40, 8512) / 1e4, rng.integers(1, 60, 8512) / 100.0])
    got = pk.KDE_JS_divergence(a, b, 1e-12, 1e-12, False, resolution=3, bandwidth=0.1)
    assert np.isfinite(got) and 0.0 <= got < 3e-8
    from oracle import workloads as wl

    x, y = wl.synthetic_target(28000), wl.sim_S(5, 2000, 0.33, n=14000)
    for gamma in (3.0, 2.0):
        got = pk.KDE_JS_divergence(x, y, gamma, 0.0, False, resolution=300)
        assert rel(got, ko.discrepancy(x, y, gamma, 0.0, False, resolution=300)) < 1e-9


def test_numa_node_query_and_bind(pk):
    import ctypes, os
    nat = pk._native
    node = ctypes.c_int(-5)
    nat.check(nat.lib().kdejs_device_numa_node(0, ctypes.byref(node)))
    assert node.value >= -1
    before = os.sched_getaffinity(0)
    try:
        got = nat.bind_host_to_device(0)
        assert got in (-1, node.value)
        assert os.sched_getaffinity(0) <= before and os.sched_getaffinity(0)
    finally:
        os.sched_setaffinity(0, before)


# ---- cell moments (cells wholly inside a tile's support enter as five moments, fine grids) ---------------
@pytest.mark.parametrize("R,cells", [(100, None), (256, None), (256, "128"), (511, None), (1024, None), (1200, "96")])
def test_cell_moments_match_the_direct_sum(pk, ko, monkeypatch, R, cells):
    from oracle import workloads as wl

    a = wl.synthetic_target(30000, seed=5)
    b = wl.sim_S(9, 2100, 0.33, n=20000)
    if cells:
        monkeypatch.setenv("KDEJS_CELLS", cells)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("KDEJS_MOMENTS", flag)
        out[flag] = pk.KDE_distance._discrepancy(a, b, 0.25, 0.0, False, R, 0.03, "epanechnikov", None, None,
                                                 want_z=True, want_density=True)
    js_o, oz1, oz2, od1, od2 = ko.discrepancy(a, b, 0.25, 0.0, resolution=R, return_grids=True)
    for flag in ("0", "1"):
        js, z1, z2, d1, d2 = out[flag]
        assert rel(js, js_o) < JS_RTOL_EXPECT
        for got, want in ((d1, od1), (d2, od2)):
            assert np.all(np.abs(got - want) <= 1e-9 * want + 1e-12 * want.max())
            assert np.array_equal(got == 0, want == 0)
    # the two paths agree far below the parity tolerance
    assert np.max(np.abs(out["0"][3] - out["1"][3])) <= 1e-12 * od1.max()


def test_cell_moments_row_sharded(pk, ko, monkeypatch):
    """With the moments on (forced: the library only enables them by itself for >= 128 points per cell), the
    row-sharded phases must agree with the single launch and with the direct sum."""
    import ctypes

    from oracle import workloads as wl

    nat = pk._native
    a = wl.synthetic_target(40000, seed=2)
    b = wl.sim_S(4, 1500, 0.4, n=40000)
    R = 1024
    monkeypatch.setenv("KDEJS_MOMENTS", "0")
    direct = pk.KDE_JS_divergence(a, b, 0.25, 0.0, resolution=R)
    monkeypatch.setenv("KDEJS_MOMENTS", "1")
    one = pk.KDE_JS_divergence(a, b, 0.25, 0.0, resolution=R)
    assert rel(one, ko.discrepancy(a, b, 0.25, 0.0, resolution=R)) < JS_RTOL_EXPECT
    assert rel(direct, one) < 1e-12       # different summation order, same number
    bounds = [0, 300, 700, R]
    common = (R, 0.03, 0.25, 0.0, 0, 0, 0)
    norms, lrs = [], []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        n = (ctypes.c_double * 2)()
        nat.check(nat.lib().kdejs_shard_phase1(0, nat.ptr(a), len(a), nat.ptr(b), len(b), *common, lo, hi, n))
        norms.append((n[0], n[1]))
    tot = (ctypes.c_double * 2)(sum(x for x, _ in norms), sum(y for _, y in norms))
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        n = (ctypes.c_double * 2)()
        lr = (ctypes.c_double * 2)()
        nat.check(nat.lib().kdejs_shard_phase1(0, nat.ptr(a), len(a), nat.ptr(b), len(b), *common, lo, hi, n))
        nat.check(nat.lib().kdejs_shard_phase2(0, tot, lr))
        lrs.append((lr[0], lr[1]))
    js = np.sqrt((sum(x for x, _ in lrs) + sum(y for _, y in lrs)) / 2.0)
    assert js == pytest.approx(one, rel=1e-12)


def test_cell_moments_switch_on_by_themselves_for_dense_cells(pk, ko):
    """1.3 M points on a 512^2 grid (64^2 cells): >= 128 points per cell, so the library takes the moment path unasked
    (profile slot 7 counts its launch) and still reproduces the oracle."""
    import ctypes

    from oracle import workloads as wl

    nat = pk._native
    L = nat.lib()
    a = wl.synthetic_target(650_000, seed=8)
    b = wl.sim_S(6, 1700, 0.37, n=650_000)
    nat.check(L.kdejs_profile_reset(0))
    nat.check(L.kdejs_profile_enable(0, 1))
    got = pk.KDE_JS_divergence(a, b, 0.25, 0.0, resolution=512)
    ms, n = ctypes.c_double(), ctypes.c_int64()
    nat.check(L.kdejs_profile_read(0, 7, ctypes.byref(ms), ctypes.byref(n)))
    nat.check(L.kdejs_profile_enable(0, 0))
    assert n.value == 1
    assert rel(got, ko.discrepancy(a, b, 0.25, 0.0, resolution=512)) < JS_RTOL_EXPECT


def test_pageable_inputs_go_through_the_staging_threads_and_change_nothing(pk):
    """Ordinary numpy arrays (what the reference's callers pass) are staged through pinned slots by worker
    threads; page-locked arrays are copied directly.  Same bits either way, for the batch, pairs and single paths,
    including ragged sizes that split into several 2 MB chunks."""
    from oracle import workloads as wl

    nat = pk._native
    rng = np.random.default_rng(3)
    obs = wl.synthetic_target(100_000)
    sizes = [100_000, 1, 131_072, 131_073, 262_145, 77, 50_000, 400_000] + [int(x) for x in rng.integers(1, 150_000, 40)]
    sims = [wl.sim_S(10 + i, 1500 + 10 * i, 0.3 + 0.002 * i, n=n) for i, n in enumerate(sizes)]
    pool = nat.PinnedPool()
    pobs = pool.empty(obs.shape); pobs[:] = obs
    psims = []
    for s in sims:
        b = pool.empty(s.shape); b[:] = s; psims.append(b)
    want = pk.KDE_JS_divergence_batch(pobs, psims, 0.25, 0.0, resolution=128)
    for _ in range(3):          # repeated calls reuse the slot ring
        got = pk.KDE_JS_divergence_batch(obs, sims, 0.25, 0.0, resolution=128)
        np.testing.assert_array_equal(got, want)
    mixed = [psims[i] if i % 3 == 0 else sims[i] for i in range(len(sims))]
    np.testing.assert_array_equal(pk.KDE_JS_divergence_batch(pobs, mixed, 0.25, 0.0, resolution=128), want)
    pairs = [(0, 1), (3, 4), (7, 4), (2, 2)]
    np.testing.assert_array_equal(pk.KDE_JS_divergence_pairs(sims[:8], pairs, gamma=0.25, eps=0.0, resolution=128),
                                  pk.KDE_JS_divergence_pairs(psims[:8], pairs, gamma=0.25, eps=0.0, resolution=128))
    big_a, big_b = wl.synthetic_target(700_001, seed=4), wl.sim_S(2, 1600, 0.35, n=333_333)
    pa = pool.empty(big_a.shape); pa[:] = big_a
    pb = pool.empty(big_b.shape); pb[:] = big_b
    assert pk.KDE_JS_divergence(big_a, big_b, 0.25, 0.0, resolution=200) == pk.KDE_JS_divergence(pa, pb, 0.25, 0.0, resolution=200)


def test_dev_entry_interleaved_with_host_calls_without_sync(pk):
    """kdejs_discrepancy_batch_dev returns with its wave still running on the caller's stream; host entry points
    (library stream) and a second caller stream reuse the same scratch right away.  The library orders them through
    an event; no result may change."""
    import ctypes

    import torch

    from oracle import workloads as wl

    nat = pk._native
    obs = wl.synthetic_target(60000)
    sims = [wl.sim_batch_member(s, n=50000) for s in range(24)]
    other = [wl.sim_batch_member(100 + s, n=7000) for s in range(6)]
    want_dev = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0, resolution=256)
    want_host = pk.KDE_JS_divergence_batch(obs[:9000], other, gamma=0.5, eps=1e-12, resolution=64)
    dobs = torch.from_numpy(obs).cuda()
    dsims = [torch.from_numpy(s).cuda() for s in sims]
    ptrs = (ctypes.c_void_p * 24)(*[t.data_ptr() for t in dsims])
    cnt = (ctypes.c_int64 * 24)(*[t.shape[0] for t in dsims])
    torch.cuda.synchronize()
    for rep in range(3):
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        out1 = torch.zeros(24, dtype=torch.float64, device="cuda")
        out2 = torch.zeros(24, dtype=torch.float64, device="cuda")
        call = lambda out, st: nat.check(nat.lib().kdejs_discrepancy_batch_dev(
            0, dobs.data_ptr(), len(obs), ptrs, cnt, 24, 256, 0.03, 0.25, 0.0, 0, 0, 0, out.data_ptr(), None,
            st.cuda_stream))
        call(out1, s1)                                               # in flight on s1 ...
        got_host = pk.KDE_JS_divergence_batch(obs[:9000], other, gamma=0.5, eps=1e-12, resolution=64)   # ... library stream
        call(out2, s2)                                               # ... and a different caller stream
        single = pk.KDE_JS_divergence(obs, sims[3], gamma=0.25, eps=0.0, resolution=256)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out1.cpu().numpy(), want_dev)
        np.testing.assert_array_equal(out2.cpu().numpy(), want_dev)
        np.testing.assert_array_equal(got_host, want_host)
        assert single == want_dev[3]


@pytest.mark.parametrize("algo", ["binned", "binned32"])
def test_very_heavy_tiles_are_processed_in_parts(pk, ko, algo):
    """Tiles with >= 2^17 candidates are split into parts that separate blocks walk; the part that finishes last
    adds the partial sums in part order.  Concentrated clouds on a coarse grid put several hundred thousand
    candidates into the central tiles."""
    rng = np.random.default_rng(17)
    a = np.vstack([rng.normal([0.5, 0.5], 0.05, (300_000, 2)), [[0.0, 0.0], [1.0, 1.0]]])
    b = np.vstack([rng.normal([0.55, 0.5], 0.07, (250_000, 2)), [[0.0, 0.0], [1.0, 1.0]]])
    for R in (20, 100):
        got = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=R, algo=algo)
        want = ko.discrepancy(a, b, 0.25, 0.0, resolution=R)
        assert rel(got, want) < (JS_RTOL_EXPECT if algo == "binned" else 2e-7)
        again = pk.KDE_JS_divergence(a, b, gamma=0.25, eps=0.0, resolution=R, algo=algo)
        assert again == got                                           # whichever block finishes last: same bits
    js, z1, z2, d1, d2 = pk.KDE_distance._discrepancy(a, b, 0.25, 0.0, False, 20, 0.03, "epanechnikov", algo, None,
                                                      want_z=True, want_density=True)
    _, _, _, od1, od2 = ko.discrepancy(a, b, 0.25, 0.0, resolution=20, return_grids=True)
    tol = 1e-12 if algo == "binned" else 1e-7
    for got_d, want_d in ((d1, od1), (d2, od2)):
        assert np.all(np.abs(got_d - want_d) <= 1e-6 * want_d + tol * want_d.max())
        assert np.array_equal(got_d == 0, want_d == 0)
