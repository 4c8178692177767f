"""CPU: host-side logic and the C-ABI surface (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _no_gpu():
    import poppunk_mod_b200 as pk

    return pk._native.device_count() == 0


def test_library_loads_and_exports_every_declared_symbol():
    import poppunk_mod_b200 as pk

    L = pk._native.lib()
    header = open(os.path.join(ROOT, "include", "kdejs.h")).read()
    declared = set(re.findall(r"\b(kdejs_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in include/kdejs.h"
    for name in declared:
        assert hasattr(L, name), f"libkdejs.so does not export {name}"
    assert declared == set(pk._native.EXPORTED_SYMBOLS)
    assert L.kdejs_version() == 100


def test_alias_and_drop_in_import_paths():
    import importlib
    import sys

    import poppunk_mod_b200 as pk

    assert callable(pk.KDE_JS_divergence) and callable(pk.KDE_KL_divergence) and callable(pk.get_grid)
    sys.path.insert(0, os.path.join(ROOT, "poppunk-mod_b200"))
    try:
        sys.modules.pop("KDE_distance", None)
        K = importlib.import_module("KDE_distance")      # `from KDE_distance import ...` as poppunk_mod.py:20 does
        for name in ("KDE_JS_divergence", "KDE_KL_divergence", "get_grid", "scale_KDE", "main"):
            assert hasattr(K, name)
    finally:
        sys.path.pop(0)
        sys.modules.pop("KDE_distance", None)


def test_signature_matches_reference():
    import inspect

    import poppunk_mod_b200 as pk

    sig = inspect.signature(pk.KDE_JS_divergence)
    names = list(sig.parameters)
    assert names[:6] == ["df1", "df2", "gamma", "eps", "log", "outplot"]        # KDE_distance.py:104
    p = sig.parameters
    assert (p["gamma"].default, p["eps"].default, p["log"].default, p["outplot"].default) == (1.0, 1e-12, False, None)
    assert p["resolution"].default == 100 and p["bandwidth"].default == 0.03    # KDE_distance.py:89, :50
    assert all(p[n].kind is inspect.Parameter.KEYWORD_ONLY for n in names[6:])


def test_get_grid_order():
    import poppunk_mod_b200 as pk

    xx, yy, xy = pk.get_grid(0, 1, 100)
    lin = np.linspace(0, 1, 100)
    assert xx.shape == yy.shape == (100, 100) and xy.shape == (10000, 2)
    k = np.arange(10000)
    np.testing.assert_array_equal(xy[:, 0], lin[k // 100])     # column 0 varies slowest
    np.testing.assert_array_equal(xy[:, 1], lin[k % 100])
    np.testing.assert_array_equal(xx[0], lin)


def test_argument_validation_happens_before_cuda():
    import poppunk_mod_b200 as pk

    good = np.random.default_rng(0).random((5, 2))
    with pytest.raises(ValueError):
        pk.KDE_JS_divergence(np.zeros((5, 3)), good)
    with pytest.raises(ValueError):
        pk.KDE_JS_divergence(np.zeros(5), good)
    with pytest.raises(ValueError, match="0 sample"):
        pk.KDE_JS_divergence(np.zeros((0, 2)), good)
    with pytest.raises(ValueError):
        pk.KDE_JS_divergence(good, good, kernel="tophat")
    with pytest.raises(ValueError):
        pk.KDE_JS_divergence(good, good, resolution=0)
    with pytest.raises(ValueError):
        pk.KDE_JS_divergence(good, good, bandwidth=-1.0)
    with pytest.raises(ValueError):
        pk.KDE_JS_divergence(good, good, kernel="gaussian", algo="binned")
    # the summary step on existing discrepancies: shape errors are raised by the wrapper, an empty batch needs no GPU
    kd = pk.KDE_distance
    with pytest.raises(ValueError, match="one gene-frequency vector"):
        kd.summary_from_js([0.5, 0.6], [[0.1]], [0, 0.3, 0.2, 0.5], 0.95, 0.05)
    with pytest.raises(ValueError, match="4-vector"):
        kd.summary_from_js([0.5], [[0.1]], [0, 0.3, 0.2], 0.95, 0.05)
    summary, d, log_d = kd.summary_from_js([], [], [0, 0.3, 0.2, 0.5], 0.95, 0.05)
    assert summary.shape == (0, 4) and d.shape == (0,) and log_d.shape == (0,)
    lib = pk._native.lib()
    assert lib.kdejs_summary_from_js(0, None, None, None, -1, 0.95, 0.05, None, None, None, None) == pk._native.KDEJS_ERR_ARG
    assert lib.kdejs_summary_from_js(0, None, None, None, 0, 0.95, 0.05, None, None, None, None) == pk._native.KDEJS_OK
    assert lib.kdejs_summary_from_js(0, None, None, None, 3, 0.95, 0.05, None, None, None, None) == pk._native.KDEJS_ERR_ARG


def test_inputs_are_not_mutated_and_views_are_accepted():
    import poppunk_mod_b200 as pk

    big = np.asfortranarray(np.random.default_rng(1).random((50, 4)))
    view = big[:, 2:]                       # non-contiguous last-two-columns slice (pairwise_2D_distance.py:45)
    pts = pk._native.as_points(view)
    assert pts.flags["C_CONTIGUOUS"] and pts.dtype == np.float64
    np.testing.assert_array_equal(pts, view)
    assert pk._native.as_points([[1, 2], [3, 4]]).dtype == np.float64


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_a_gpu():
    import poppunk_mod_b200 as pk

    a = np.random.default_rng(0).random((20, 2))
    with pytest.raises(RuntimeError, match="kdejs"):
        pk.KDE_JS_divergence(a, a)
    with pytest.raises(RuntimeError, match="kdejs"):
        pk.KDE_JS_divergence_batch(a, [a, a])
    with pytest.raises(RuntimeError, match="kdejs"):
        pk.kde_density(a)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "poppunk-mod_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"


def test_default_device_env(monkeypatch):
    import poppunk_mod_b200 as pk

    monkeypatch.delenv("KDEJS_DEVICE", raising=False)
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    assert pk.default_device() == 0
    monkeypatch.setenv("KDEJS_DEVICE", "3")
    assert pk.default_device() in (3, 3 % max(1, pk._native.device_count()))


def test_cpulist_parser_and_numa_bind_opt_out(monkeypatch):
    import os
    import poppunk_mod_b200 as pk

    nat = pk._native
    assert nat._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert nat._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    monkeypatch.setenv("KDEJS_NUMA_BIND", "0")
    assert nat.bind_host_to_device(0) == -1
    monkeypatch.delenv("KDEJS_NUMA_BIND")
    # without a GPU the node query fails: nothing may change
    if not nat.device_count():
        assert nat.bind_host_to_device(0) == -1
        assert os.sched_getaffinity(0) == before


def test_default_algorithm_and_its_override():
    """The package default is the fp32 culled kernel; KDEJS_DEFAULT_ALGO selects another one for a whole process."""
    import subprocess
    import sys

    code = "import sys; sys.path.insert(0, %r); import poppunk_mod_b200 as pk; print(pk.KDE_distance.DEFAULT_ALGO)" % ROOT
    env = {k: v for k, v in os.environ.items() if k != "KDEJS_DEFAULT_ALGO"}
    assert subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env).stdout.strip() == "binned32"
    env["KDEJS_DEFAULT_ALGO"] = "binned"
    assert subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env).stdout.strip() == "binned"
    import poppunk_mod_b200 as pk

    assert set(pk._native.ALGOS) == {"binned", "binned32", "dense", "dense64"}
    with pytest.raises(ValueError):
        pk.KDE_distance._common(100, 0.03, 1.0, 0.0, False, "epanechnikov", "fastest")


def test_point_sets_list_and_3d_array_agree():
    """PointSets: addresses and row counts of a batch without touching the device -- a (B,N,2) array and the list of
    its rows give the same table; views that are not C-contiguous (N,2) float64 are copied, never mutated."""
    import poppunk_mod_b200 as pk

    nat = pk._native
    arena = np.random.default_rng(0).random((9, 50, 2))
    a, b = nat.PointSets([arena[i] for i in range(9)]), nat.PointSets(arena)
    assert len(a) == len(b) == 9
    np.testing.assert_array_equal(a.addr, b.addr)
    np.testing.assert_array_equal(a.count, b.count)
    np.testing.assert_array_equal(nat.PointSets(arena[::2]).addr, b.addr[::2])
    ro = arena.copy()
    ro.flags.writeable = False
    c = nat.PointSets([ro[i] for i in range(9)])                       # read-only arrays: the slower address path
    np.testing.assert_array_equal(c.addr - c.addr[0], a.addr - a.addr[0])
    wide = np.arange(40.0).reshape(10, 4)
    v = nat.PointSets([wide[:, :2], wide[:, 1:3], [[1, 2], [3, 4]], np.ones((3, 2), dtype=np.float32)])
    assert all(k.flags.c_contiguous and k.dtype == np.float64 for k in v.keep)
    assert v.count.tolist() == [10, 10, 2, 3] and v.sizes() == {2, 3, 10}
    np.testing.assert_array_equal(v.keep[1], wide[:, 1:3])
    f32 = nat.PointSets(arena.astype(np.float32))                      # wrong dtype: one contiguous float64 copy
    assert f32.keep.dtype == np.float64 and f32.addr[1] - f32.addr[0] == 50 * 16
    ptrs, counts, n, keep = b.pointers(slice(1, 9, 3))
    assert n == 3 and keep[0].tolist() == b.addr[1:9:3].tolist()
    for bad in (np.zeros((3, 0, 2)), np.zeros((3, 5, 3))):
        with pytest.raises(ValueError):
            nat.PointSets(bad)
    with pytest.raises(ValueError):
        nat.PointSets([np.zeros((0, 2))])


def test_sweep_arenas_are_leased_not_shared():
    """The page-locked arenas the sweep drivers keep between calls: reused by the next sweep, grown when a longer or
    larger one comes, never handed to two sweeps at once, given back by release_arenas()."""
    from poppunk_mod_b200 import sweeps

    class Arena:
        made = 0

        def __init__(self, slots, cap):
            Arena.made += 1
            self.cap, self.closed = cap, False

        def close(self):
            self.closed = True

    make = lambda n, c: Arena(n, c)
    with sweeps._ArenaLease("test", 1, 4, 100, make) as first:
        with sweeps._ArenaLease("test", 2, 4, 100, make) as second:            # the key is busy: arenas of its own
            assert len(second) == 2 and second[0] is not first[0]
        assert all(a.closed for a in second) and not first[0].closed
    with sweeps._ArenaLease("test", 2, 4, 50, make) as again:                   # reused, one arena added
        assert again[0] is first[0] and len(again) == 2
    with sweeps._ArenaLease("test", 1, 4, 500, make) as larger:                 # too small: replaced
        assert larger[0] is not first[0] and first[0].closed and larger[0].cap == 500
    sweeps.release_arenas()
    assert larger[0].closed and ("test", 4) not in sweeps._ARENAS and Arena.made == 5


def test_simulator_batch_checks_and_table_shape_detection(tmp_path):
    """simulator.process_results_batch: argument errors surface before any CUDA work; the device parser is only
    chosen for tables read_distfile (poppunk_mod.py:255-269) would read whole -- two tab-separated columns."""
    from poppunk_mod_b200 import simulator

    obs = np.random.default_rng(0).random((10, 2))
    with pytest.raises(ValueError, match="parse"):
        simulator.process_results_batch([obs], [[0.1]], obs, 0.25, 0.95, 0.05, parse="gpu")
    with pytest.raises(ValueError, match="one gene-frequency vector per simulation"):
        simulator.process_results_batch([obs, obs], [[0.1]], obs, 0.25, 0.95, 0.05)
    two, four, trailing = tmp_path / "two.tsv", tmp_path / "four.tsv", tmp_path / "trailing.tsv"
    two.write_text("0.001\t0.2\n0.002\t0.3\n")
    four.write_text("A\tB\t0.001\t0.2\n")
    trailing.write_text("0.001\t0.2\t\r\n")                # str.rstrip() drops the trailing tab, as in read_distfile
    assert simulator._two_column(two) and simulator._two_column(trailing) and not simulator._two_column(four)
    many = [str(two)] * 70
    assert simulator._all_two_column(many) and not simulator._all_two_column(many + [str(four)])
    assert not simulator._all_two_column([str(four), str(two)])
