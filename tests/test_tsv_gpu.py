"""GPU: tables parsed ON THE DEVICE (csrc/kdejs_tsv.cuh, kdejs_decimal.h) against the host reader / np.loadtxt / float():
the same rows bit for bit, the same discrepancies as the host path, and a clean hand-over of everything the device parser
does not cover."""
import math
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pk():
    import poppunk_mod_b200 as pk

    if pk._native.device_count() < 1:
        pytest.skip("no CUDA device")
    return pk


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _table_text(a, ids=False, crlf=False):
    rows = []
    for k, (x, y) in enumerate(np.asarray(a).tolist()):
        f = ([f"S{k}", f"T{k * 7}"] if ids else []) + [repr(x), repr(y)]
        rows.append("\t".join(f) + ("\r" if crlf else ""))
    return ("\n".join(rows) + "\n").encode()


def test_device_rows_equal_the_host_reader(pk, tmp_path):
    nat = pk._native
    rng = np.random.default_rng(0)
    a = np.column_stack([rng.integers(0, 2000, 50000) / 1e6, rng.random(50000)])
    for kw in (dict(), dict(ids=True), dict(crlf=True)):
        text = _table_text(a, **kw)
        got, st = nat.tsv_parse_text(text, skip_rows=0)
        assert st == nat.KDEJS_TABLE_OK
        np.testing.assert_array_equal(_bits(got), _bits(a))
        p = tmp_path / "t.tsv"
        p.write_bytes(text)
        np.testing.assert_array_equal(_bits(got), _bits(nat.TsvTable(str(p), skip_rows=0, pinned=False).array))
    # header rules: auto (-1) drops a non-numeric first row only; explicit counts drop data rows
    text = b"core\taccessory\n" + _table_text(a[:100])
    got, st = nat.tsv_parse_text(text, skip_rows=-1)
    assert st == 0 and got.shape == (100, 2)
    got, _ = nat.tsv_parse_text(_table_text(a[:100]), skip_rows=-1)
    assert got.shape == (100, 2)
    got, _ = nat.tsv_parse_text(_table_text(a[:100]), skip_rows=7)
    np.testing.assert_array_equal(got, a[7:100])
    # no newline at the end; a file that is one line; blank lines and comments only
    got, st = nat.tsv_parse_text(_table_text(a[:3])[:-1], skip_rows=0)
    assert st == 0
    np.testing.assert_array_equal(got, a[:3])
    got, st = nat.tsv_parse_text(b"0.5\t0.25", skip_rows=0)
    assert st == 0 and got.tolist() == [[0.5, 0.25]]
    got, st = nat.tsv_parse_text(b"\n# nothing\n\n", skip_rows=0)
    assert got.shape == (0, 2) and st == nat.KDEJS_TABLE_HOST


def test_ragged_text_lands_every_row_where_the_host_puts_it(pk, tmp_path):
    """Blank lines, comments, CRLF, id columns of very different length (lines longer and shorter than a thread's 32
    bytes, lines across the 8 KB chunks of a block), blanks around fields, an empty field (NaN)."""
    nat = pk._native
    rng = np.random.default_rng(3)
    lines = []
    for i in range(80000):
        r = rng.random()
        if r < 0.02:
            lines.append("")
        elif r < 0.04:
            lines.append("# comment " + "x" * int(rng.integers(0, 300)))
        elif r < 0.05:
            lines.append("a\tb\t\t0.5")                       # empty field: missing value
        else:
            x, y = float(rng.random() * 1e-3), float(rng.random())
            ids = ["id" * int(rng.integers(1, 60)), "other"] if r < 0.5 else []
            lines.append("\t".join(ids + [repr(x), f" {y!r} "]) + ("\r" if r > 0.9 else ""))
    text = "\n".join(lines).encode()
    p = tmp_path / "ragged.tsv"
    p.write_bytes(text)
    want = nat.TsvTable(str(p), skip_rows=0, pinned=False, threads=1).array
    got, st = nat.tsv_parse_text(text, skip_rows=0)
    assert st == nat.KDEJS_TABLE_OK and got.shape == want.shape
    np.testing.assert_array_equal(_bits(got), _bits(want))
    got5, _ = nat.tsv_parse_text(text, skip_rows=5)
    np.testing.assert_array_equal(_bits(got5), _bits(want[5:]))


def test_device_decimal_conversion_is_correctly_rounded(pk):
    """kdejs_decimal.h on the device against float(): spellings of every length, decimals next to the midpoint of two
    doubles, exact ties; spellings outside its range flag the table instead of producing a value."""
    from decimal import ROUND_CEILING, ROUND_FLOOR, Decimal, localcontext

    nat = pk._native
    rng = np.random.default_rng(11)
    random.seed(5)
    strs = []
    fmts = ["%r", "%.17g", "%.18e", "%.6f", "%.15g", "%.16g", "%.19g", "%.3f"]
    vals = rng.random(120000) * 10.0 ** rng.uniform(-12, 12, 120000) * np.where(rng.random(120000) < 0.3, -1, 1)
    for i, v in enumerate(vals.tolist()):
        f = fmts[i % len(fmts)]
        strs.append(repr(v) if f == "%r" else f % v)
    for _ in range(60000):
        nd = random.randint(1, 19)
        digs = "".join(random.choice("0123456789") for _ in range(nd))
        pt = random.randint(0, nd)
        s = (digs[:pt] or "0") + "." + digs[pt:]
        if random.random() < 0.3:
            s += random.choice("eE") + random.choice(["", "+", "-"]) + str(random.randint(0, 8))
        strs.append(("-" if random.random() < 0.2 else "+" if random.random() < 0.05 else "") + s)
    for v in (rng.random(30000) * 10.0 ** rng.uniform(-8, 8, 30000)).tolist():
        with localcontext() as ctx:
            ctx.prec = 60
            mid = (Decimal(v) + Decimal(math.nextafter(v, math.inf))) / 2
        for rounding in (ROUND_FLOOR, ROUND_CEILING):
            with localcontext() as ctx:
                ctx.prec, ctx.rounding = 19, rounding
                strs.append(format(+mid, "e"))
    for _ in range(20000):
        base = random.randrange(2 ** 53, 2 ** 62)
        strs += [str(base), f"{base}.5"][: 2 if len(str(base)) < 19 else 1]
    strs += ["0", "-0.0", "0.000", "1.", ".5", "5e3", "1E+2", "00012.50", "9007199254740993", "1e27", "1e-27"]

    def covered(s):            # the range kdejs_decimal.h handles: <= 19 significant digits, |q| <= 27
        m, _, e = s.lower().lstrip("+-").partition("e")
        ip, _, fp = m.partition(".")
        sig = len((ip + fp).lstrip("0"))
        q = int(e or 0) - len(fp)
        return sig <= 19 and (sig == 0 or -27 <= q <= 27)

    n_all = len(strs)
    strs = [s for s in strs if covered(s)]
    assert len(strs) > 0.9 * n_all
    if len(strs) % 2:
        strs.append("1")
    want = np.array([float(s) for s in strs]).reshape(-1, 2)
    text = "".join(f"{strs[i]}\t{strs[i + 1]}\n" for i in range(0, len(strs), 2)).encode()
    got, st = nat.tsv_parse_text(text, skip_rows=0)
    assert st == nat.KDEJS_TABLE_OK
    bad = np.argwhere(_bits(got) != _bits(want))
    assert len(bad) == 0, [(strs[2 * r + c], got[r, c], want[r, c]) for r, c in bad[:5]]
    for outside in ("NA", "nan", "inf", "1e400", "1e-30", "12345678901234567890", "0x10", "1e", "--1", "1 2", "abc"):
        _, st = nat.tsv_parse_text(f"0.1\t0.2\n0.3\t{outside}\n".encode(), skip_rows=0)
        assert st == nat.KDEJS_TABLE_HOST, outside
    _, st = nat.tsv_parse_text(b"0.1\t0.2\njustonefield\n", skip_rows=0)
    assert st == nat.KDEJS_TABLE_HOST


def test_batch_text_equals_the_host_path(pk, tmp_path):
    """kdejs_discrepancy_batch_text: discrepancies, row counts and bounds of 37 tables of different length (three
    waves) equal parsing on the host and calling the batch entry point; tables the device parser does not cover come
    back flagged, not wrong."""
    from oracle import workloads as wl

    nat = pk._native
    obs = wl.synthetic_target(20000)
    sims = [wl.sim_batch_member(s, n=3000 + 811 * s) for s in range(37)]
    texts = [_table_text(s, ids=(k % 3 == 0)) for k, s in enumerate(sims)]
    texts[5] = texts[5].replace(b"\n", b"\nS\tT\tNA\t0.5\n", 1)            # a missing-value spelling: host reader
    texts[9] = b"# empty\n"                                                # no rows
    texts[11] = texts[11].replace(b"\n", b"\nS\tT\t\t0.5\n", 1)           # an empty field: NaN, handled on the device
    sims[11] = np.vstack([sims[11][:1], [[np.nan, 0.5]], sims[11][1:]])
    want = pk.KDE_JS_divergence_batch(obs, sims, gamma=0.25, eps=0.0)
    js, rows, status, bounds = pk.KDE_JS_divergence_batch_text(obs, texts, gamma=0.25, eps=0.0, skip_rows=0, bounds=True)
    ok = [i for i in range(37) if i not in (5, 9)]
    assert status[ok].tolist() == [nat.KDEJS_TABLE_OK] * len(ok)
    assert status[5] == nat.KDEJS_TABLE_HOST and status[9] == nat.KDEJS_TABLE_HOST
    assert np.isnan(js[5]) and np.isnan(js[9]) and rows[9] == 0
    np.testing.assert_array_equal(js[ok], want[ok])                        # same doubles in, same kernels: same bits
    assert rows[ok].tolist() == [sims[i].shape[0] for i in ok]
    for i in ok:
        s = sims[i]
        np.testing.assert_array_equal(bounds[i], [np.min(s[:, 0]), np.max(s[:, 0]), np.min(s[:, 1]), np.max(s[:, 1])])
    # the same tables as rows of one page-locked byte arena read from disk
    paths = []
    for k, t in enumerate(texts):
        q = tmp_path / f"t{k}.tsv"
        q.write_bytes(t)
        paths.append(str(q))
    arena = nat.TextArena(37, max(len(t) for t in texts) + 100)
    n = arena.read(paths + [], threads=4)
    assert n.tolist() == [len(t) for t in texts]
    js2, rows2, status2 = pk.KDE_JS_divergence_batch_text(obs, arena.array, gamma=0.25, eps=0.0, skip_rows=0, nbytes=n)
    np.testing.assert_array_equal(js2[ok], want[ok])
    np.testing.assert_array_equal(status2, status)
    small = nat.TextArena(2, 1000)
    assert small.read(paths[:1] + [str(tmp_path / "missing.tsv")]).tolist() == [-len(texts[0]), -1]
    arena.close()
    small.close()


def test_gridsearch_device_and_host_parsing_agree(pk, tmp_path):
    """sweeps.gridsearch_best with the files' bytes parsed on the device against the host-parser pipeline: the same
    distances bit for bit, the same best file and bounds -- including a table only the host reader covers (NA), one
    too long for its slot, varying lengths, and the bounds filter of distance_to_gridsearch.py:112-125."""
    from oracle import workloads as wl
    from poppunk_mod_b200 import sweeps

    ref = wl.synthetic_target(15000)
    sims = [wl.sim_batch_member(s, n=4000 + 503 * (s % 7)) for s in range(23)]
    sims[13] = wl.sim_batch_member(13, n=30000)                                # far longer than 1.5 x the first table
    texts = [_table_text(s) for s in sims]
    texts[4] = texts[4].replace(b"\n", b"\nNA\t0.5\n", 1)
    sims[4] = np.vstack([sims[4][:1], [[np.nan, 0.5]], sims[4][1:]])
    paths = []
    for k, t in enumerate(texts):
        q = tmp_path / f"sim{k}.tsv"
        q.write_bytes(t)
        paths.append(str(q))
    refp = tmp_path / "ref.tsv"
    refp.write_bytes(_table_text(ref))
    kw = dict(gamma=0.25, eps=0.0, log=False, chunk=8, io_threads=4)
    best_d, dist_d = sweeps.gridsearch_best(str(refp), paths, parse="device", **kw)
    best_h, dist_h = sweeps.gridsearch_best(str(refp), paths, parse="host", **kw)
    np.testing.assert_array_equal(dist_d, dist_h)
    assert best_d[0] == best_h[0] and best_d[1] == best_h[1]
    np.testing.assert_array_equal(np.array(best_d[2]), np.array(best_h[2]))
    want = pk.KDE_JS_divergence_batch(ref, sims, gamma=0.25, eps=0.0)
    np.testing.assert_array_equal(dist_d, want)
    # a bounds filter that excludes the overall best: both pipelines keep the same runner-up
    order = np.argsort(want)
    b0 = sims[order[0]]
    lo = float(np.nanmax(b0[:, 1])) * 1.0000001
    kwb = dict(acc_min_bounds=[0.0, 1.0], acc_max_bounds=[lo, 10.0]) if all(
        np.nanmax(s[:, 1]) != np.nanmax(b0[:, 1]) for k, s in enumerate(sims) if k != order[0]) else {}
    bd, _ = sweeps.gridsearch_best(str(refp), paths, parse="device", **kw, **kwb)
    bh, _ = sweeps.gridsearch_best(str(refp), paths, parse="host", **kw, **kwb)
    assert bd[0] == bh[0] and bd[1] == bh[1]


def test_degenerate_tables_do_not_break_the_device_parser(pk):
    """Thousands of two-byte lines in one 8 KB chunk (the most line starts a chunk can hold), lines much longer than the
    window staged in shared memory, a table that is all tabs: flagged or parsed like the host reader, never out of bounds."""
    nat = pk._native
    _, st = nat.tsv_parse_text(b"x\n" * 20000, skip_rows=0)                    # rows without two fields
    assert st == nat.KDEJS_TABLE_HOST
    long_id = "L" * 5000                                                        # lines of 5 KB: across chunks and windows
    text = "".join(f"{long_id}\tid\t{k * 0.5!r}\t{k * 0.25!r}\n" for k in range(300)).encode()
    got, st = nat.tsv_parse_text(text, skip_rows=0)
    assert st == nat.KDEJS_TABLE_OK
    np.testing.assert_array_equal(got, [[k * 0.5, k * 0.25] for k in range(300)])
    got, st = nat.tsv_parse_text(b"\t\t\t\n" * 5000 + b"1\t2\n", skip_rows=0)   # blank lines (tabs only) are skipped
    assert st == nat.KDEJS_TABLE_OK and got.tolist() == [[1.0, 2.0]]
    got, st = nat.tsv_parse_text(b"1\t2\n" * 70000, skip_rows=0)               # four-byte rows: 2 048 per chunk
    assert st == nat.KDEJS_TABLE_OK and got.shape == (70000, 2) and (got == [1.0, 2.0]).all()


def test_gridsearch_over_several_devices_equals_one(pk, tmp_path):
    """Chunks dealt to the visible devices round-robin (one host thread per device): the same distances and the same
    best file as on one device."""
    from oracle import workloads as wl
    from poppunk_mod_b200 import sweeps

    n_dev = pk._native.device_count()
    if n_dev < 2:
        pytest.skip("needs two GPUs")
    ref = wl.synthetic_target(12000)
    paths = []
    for k in range(29):
        q = tmp_path / f"sim{k}.tsv"
        q.write_bytes(_table_text(wl.sim_batch_member(k, n=3000 + 97 * k)))
        paths.append(str(q))
    kw = dict(gamma=0.25, eps=0.0, log=False, chunk=4, io_threads=4)
    best1, dist1 = sweeps.gridsearch_best(ref, paths, devices=0, **kw)
    bestn, distn = sweeps.gridsearch_best(ref, paths, devices=list(range(n_dev)), **kw)
    np.testing.assert_array_equal(distn, dist1)
    assert bestn[0] == best1[0] and bestn[1] == best1[1]


def test_simulator_batch_from_files_device_parser_equals_host_reader(pk, tmp_path):
    """simulator.process_results_batch over files (process_result's loads, poppunk_mod.py:360-361): with the tables
    parsed on the device and the summary kernel run on the resulting discrepancies (kdejs_summary_from_js) the vectors
    and both ELFI nodes equal the host-parsed kdejs_summary_batch path bit for bit -- including a table the device
    parser hands back to the host reader and more than one chunk of files."""
    from oracle import workloads as wl
    from poppunk_mod_b200 import simulator

    rng = np.random.default_rng(5)
    obs = wl.synthetic_target(6000)
    B = 9
    sim_paths, freq_paths, freqs = [], [], []
    for i in range(B):
        a = wl.sim_batch_member(i, n=3000 + 211 * i)
        p = tmp_path / f"sim_{i}.tsv"
        np.savetxt(p, a, delimiter="\t", fmt="%.17g" if i % 2 else "%.6f")
        if i == 4:                                  # a missing value: only the host reader covers the spelling
            lines = p.read_text().split("\n")
            lines[7] = lines[7].split("\t")[0] + "\tNA"
            p.write_text("\n".join(lines))
        f = np.round(rng.beta(0.3, 0.3, 2000 + 37 * i), 3) * (rng.random(2000 + 37 * i) > 0.2)
        fp = tmp_path / f"sim_{i}_freqs.txt"
        np.savetxt(fp, f, fmt="%.3f")
        sim_paths.append(str(p))
        freq_paths.append(str(fp))
        freqs.append(np.loadtxt(fp))
    observed = np.array([0.0, 0.3, 0.2, 0.5])
    args = (obs, 0.25, 0.95, 0.05)
    host = simulator.process_results_batch(sim_paths, freq_paths, *args, observed=observed, parse="host")
    dev = simulator.process_results_batch(sim_paths, freq_paths, *args, observed=observed, parse="device")
    for got, want in zip(dev, host):
        assert np.isfinite(want).all()
        np.testing.assert_array_equal(_bits(got), _bits(want))
    # several chunks of files, the last one short
    chunked = simulator._from_files_on_device(obs, sim_paths, freq_paths, 0.25, 0.95, 0.05, observed, None, 100, chunk=4)
    for got, want in zip(chunked, host):
        np.testing.assert_array_equal(_bits(got), _bits(want))
    # the summary step on its own: the discrepancies of the array path in, kdejs_summary_batch's vectors out
    arrays = [np.loadtxt(p, delimiter="\t") for k, p in enumerate(sim_paths) if k != 4]
    fr = [f for k, f in enumerate(freqs) if k != 4]
    s_ref, d_ref, l_ref = pk.KDE_distance.summary_batch(obs, arrays, fr, observed, 0.25, 0.95, 0.05)
    s_js, d_js, l_js = pk.KDE_distance.summary_from_js(s_ref[:, 0], fr, observed, 0.95, 0.05)
    for got, want in ((s_js, s_ref), (d_js, d_ref), (l_js, l_ref)):
        np.testing.assert_array_equal(_bits(got), _bits(want))
    keep = [k for k in range(B) if k != 4]
    np.testing.assert_array_equal(_bits(s_ref), _bits(host[0][keep]))
    # without `observed` only the vectors come back; one file alone takes the host reader
    only = simulator.process_results_batch(sim_paths[:3], freq_paths[:3], *args)
    np.testing.assert_array_equal(_bits(only), _bits(host[0][:3]))
    one = simulator.process_result(sim_paths[1], freq_paths[1], *args)
    np.testing.assert_array_equal(_bits(one), _bits(host[0][1]))
    with pytest.raises(ValueError, match="parse"):
        simulator.process_results_batch(sim_paths, freq_paths, *args, parse="gpu")
