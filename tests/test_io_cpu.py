"""CPU: the C++ table reader against numpy / pandas, and the sweep drivers' host-side logic."""
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def nat():
    import poppunk_mod_b200 as pk

    return pk._native


def _write(path, rows, header=None):
    with open(path, "w") as f:
        if header:
            f.write(header + "\n")
        for r in rows:
            f.write("\t".join(r) + "\n")


def test_two_column_matches_loadtxt(nat, tmp_path):
    rng = np.random.default_rng(0)
    a = np.column_stack([rng.integers(0, 2000, 20000) / 1e6, rng.random(20000)])
    a[17] = [0.0, 1.0]
    a[18, 1] = np.nan
    p = tmp_path / "two.tsv"
    _write(p, [[repr(float(x)), repr(float(y))] for x, y in a])
    t = nat.TsvTable(str(p), pinned=False)
    ref = np.loadtxt(p, delimiter="\t", dtype="float64")
    np.testing.assert_array_equal(t.array, ref)            # bit-identical parsing, NaN included
    assert t.array.shape == (20000, 2)
    t.close()


def test_four_column_last_two_and_header_rules(nat, tmp_path):
    import pandas as pd
    from poppunk_mod_b200 import sweeps

    rows = [[f"S{i}", f"T{i}", f"{i * 1.25e-5:.8g}", f"{0.1 + i * 1e-3:.7g}"] for i in range(5000)]
    p = tmp_path / "four.txt"
    _write(p, rows)
    df = pd.read_csv(p, header=None, sep="\t", index_col=False)
    want = df[df.columns[[-2, -1]]].to_numpy()
    np.testing.assert_array_equal(sweeps.read_file(str(p), pinned=False).array, want)     # pairwise_2D_distance.py:38-45
    # poppunk_mod.read_distfile: pandas' inferred header eats the first row of a 4-column table
    want2 = pd.read_csv(p, sep="\t").to_numpy()[:, 2:].astype(float)
    np.testing.assert_array_equal(sweeps.read_distfile(str(p), pinned=False).array, want2)
    h = tmp_path / "hdr.tsv"
    _write(h, [["1e-3", "0.5"], ["2e-3", "+0.25"]], header="core\taccessory")
    np.testing.assert_array_equal(nat.TsvTable(str(h), pinned=False).array, [[1e-3, 0.5], [2e-3, 0.25]])


def test_reader_errors(nat, tmp_path):
    with pytest.raises(ValueError, match="cannot open"):
        nat.TsvTable(str(tmp_path / "missing.tsv"), pinned=False)
    bad = tmp_path / "bad.tsv"
    _write(bad, [["1.0", "2.0"], ["x", "y"]])
    with pytest.raises(ValueError, match="row 2"):
        nat.TsvTable(str(bad), pinned=False)
    empty = tmp_path / "empty.tsv"
    empty.write_text("\n\n")
    with pytest.raises(ValueError, match="no data rows"):
        nat.TsvTable(str(empty), pinned=False)


def test_missing_values_become_nan_like_pandas(nat, tmp_path):
    """pd.read_csv (scripts/pairwise_2D_distance.py:38-45, poppunk_mod.py:266-267) turns NA tokens and empty fields
    into NaN, and the path maps NaN to scaled 0 (KDE_distance.py:53): tables with missing distances must load."""
    import pandas as pd

    f = tmp_path / "missing.tsv"
    f.write_text("a\tb\t0.001\t0.2\n"
                 "a\tc\tNA\t0.3\n"
                 "a\td\t0.002\t\n"
                 "b\tc\tnan\tN/A\n"
                 "b\td\t#N/A\tnull\n"
                 "c\td\t0.004\t0.5\n")
    got = nat.TsvTable(str(f), skip_rows=0, pinned=False).array
    want = pd.read_csv(str(f), header=None, sep="\t").iloc[:, -2:].to_numpy(dtype=np.float64)
    np.testing.assert_array_equal(got, want)                   # NaN positions and values alike
    assert np.isnan(got).sum() == 6
    junk = tmp_path / "junk.tsv"
    junk.write_text("0.1\t0.2\n0.3\tabc\n")
    with pytest.raises(ValueError, match="row 2"):
        nat.TsvTable(str(junk), pinned=False)


@pytest.mark.skipif(not os.path.exists("/root/reference/distances/2_column/Escherichia_coli.txt"),
                    reason="reference fixtures not present")
def test_reference_fixture_files(nat):
    f = "/root/reference/distances/2_column/Escherichia_coli.txt"
    np.testing.assert_array_equal(nat.TsvTable(f, pinned=False).array, np.loadtxt(f, delimiter="\t"))


def test_log_discrepancy_of_host_summaries():
    from poppunk_mod_b200 import simulator as sim

    s = np.array([[0.3, 0.5, 0.2, 0.3], [0.0, 0.4, 0.3, 0.3]])
    o = np.array([0.0, 0.4, 0.3, 0.3])
    with np.errstate(divide="ignore"):
        d = sim.log_euclidean_discrepancy(s, o)
    assert d[0] == pytest.approx(np.log(np.sqrt(0.09 + 0.01 + 0.01)))
    assert d[1] == -np.inf


def test_threaded_ranges_agree_with_one_thread_on_ragged_tables(nat, tmp_path):
    """The file is cut into byte ranges, one per thread: blank lines, comments, CRLF endings, lines of very
    different length and a missing final newline must land every row where one thread puts it."""
    rng = np.random.default_rng(3)
    lines, want = [], []
    for i in range(60000):
        r = rng.random()
        if r < 0.02:
            lines.append("")
        elif r < 0.04:
            lines.append("# comment " + "x" * int(rng.integers(0, 300)))
        else:
            x, y = float(rng.random() * 1e-3), float(rng.random())
            ids = ["id" * int(rng.integers(1, 40)), "other"] if r < 0.5 else []
            lines.append("\t".join(ids + [repr(x), f" {y!r} "]) + ("\r" if r > 0.9 else ""))
            want.append((x, y))
    p = tmp_path / "ragged.tsv"
    p.write_text("\n".join(lines))                       # no newline at the end of the file
    want = np.array(want)
    for threads in (1, 2, 3, 8):
        t = nat.TsvTable(str(p), skip_rows=0, pinned=False, threads=threads)
        np.testing.assert_array_equal(t.array, want)
    t = nat.TsvTable(str(p), skip_rows=5, pinned=False, threads=4)
    np.testing.assert_array_equal(t.array, want[5:])


def test_load_into_and_table_arena(nat, tmp_path):
    """kdejs_tsv_load_into: the same rows as kdejs_tsv_load in the caller's buffer; a buffer that is too small is
    left alone and the rows needed are reported.  TableArena: many files into one (slots, capacity, 2) array."""
    rng = np.random.default_rng(5)
    tabs, paths = [], []
    for k, n in enumerate((3000, 3000, 3000, 4100, 17)):
        a = np.column_stack([rng.random(n) * 2e-3, rng.random(n)])
        q = tmp_path / f"s{k}.tsv"
        _write(q, [[repr(float(x)), repr(float(y))] for x, y in a])
        tabs.append(a)
        paths.append(str(q))
    dst = np.full((3500, 2), -1.0)
    assert nat.tsv_load_into(paths[0], dst) == 3000
    np.testing.assert_array_equal(dst[:3000], tabs[0])
    assert (dst[3000:] == -1.0).all()
    dst[:] = -1.0
    assert nat.tsv_load_into(paths[3], dst) == -4100 and (dst == -1.0).all()
    with pytest.raises(ValueError):
        nat.tsv_load_into(paths[0], np.zeros((10, 3)))
    with pytest.raises(ValueError, match="cannot open"):
        nat.tsv_load_into(str(tmp_path / "none.tsv"), dst)
    arena = nat.TableArena(5, 3500, pinned=False)
    same = arena.load(paths[:3], workers=3)
    assert isinstance(same, np.ndarray) and same.shape == (3, 3000, 2)      # equal lengths: one 3-D view
    for k in range(3):
        np.testing.assert_array_equal(same[k], tabs[k])
    with pytest.raises(ValueError, match="do not fit"):
        arena.load(paths, workers=2)
    mixed = arena.load(paths, workers=2, overflow="alloc")                  # the long table gets memory of its own
    assert isinstance(mixed, list) and [m.shape[0] for m in mixed] == [3000, 3000, 3000, 4100, 17]
    for m, a in zip(mixed, tabs):
        np.testing.assert_array_equal(m, a)
    sets = nat.PointSets(same)
    assert sets.addr[1] - sets.addr[0] == 3500 * 16 and sets.count.tolist() == [3000] * 3
    arena.close()


def test_decimal_conversion_is_correctly_rounded(nat, tmp_path):
    """The reader's numbers against float() (what np.loadtxt and pandas produce): bit for bit, on the spellings tables
    hold, on digit strings of every length, on decimals right next to the midpoint of two doubles, and out of range
    (1e400 -> inf, 1e-400 -> 0: from_chars reports those as errors, the reader must not)."""
    import math
    import random
    from decimal import ROUND_CEILING, ROUND_FLOOR, Decimal, localcontext

    rng = np.random.default_rng(11)
    random.seed(5)
    strs = []
    fmts = ["%r", "%.17g", "%.18e", "%.6f", "%.15g", "%.16g", "%.19g", "%.20g", "%.3f"]
    vals = rng.random(40000) * 10.0 ** rng.uniform(-30, 30, 40000) * np.where(rng.random(40000) < 0.3, -1, 1)
    for i, v in enumerate(vals.tolist()):
        f = fmts[i % len(fmts)]
        strs.append(repr(v) if f == "%r" else f % v)
    strs += [repr(float(x)) for x in rng.random(20000).tolist()]
    for _ in range(20000):
        nd = random.randint(1, 20)
        digs = "".join(random.choice("0123456789") for _ in range(nd))
        pt = random.randint(0, nd)
        s = digs[:pt] + "." + digs[pt:]
        if s == ".":
            s = "0."
        if random.random() < 0.3:
            s += random.choice("eE") + random.choice(["", "+", "-"]) + str(random.randint(0, 40))
        strs.append(("-" if random.random() < 0.2 else "+" if random.random() < 0.05 else "") + s)
    for v in (rng.random(10000) * 10.0 ** rng.uniform(-8, 8, 10000)).tolist():
        with localcontext() as ctx:
            ctx.prec = 60
            mid = (Decimal(v) + Decimal(math.nextafter(v, math.inf))) / 2
        for rounding in (ROUND_FLOOR, ROUND_CEILING):
            with localcontext() as ctx:
                ctx.prec, ctx.rounding = 19, rounding
                strs.append(format(+mid, "e"))
    strs += ["0", "-0.0", "0.000", "1.", ".5", "5e3", "1E+2", "00012.50", "inf", "-inf", "nan", "1e400", "1e-400",
             "18446744073709551615", "18446744073709551616", "9007199254740993", "4.9e-324"]
    if len(strs) % 2:
        strs.append("1")
    want = np.array([float(s) for s in strs]).reshape(-1, 2)
    p = tmp_path / "numbers.tsv"
    with open(p, "w") as f:
        for i in range(0, len(strs), 2):
            f.write(strs[i] + "\t" + strs[i + 1] + "\n")
    got = nat.TsvTable(str(p), skip_rows=0, pinned=False, threads=3).array
    bad = np.argwhere(got.view(np.uint64) != want.view(np.uint64))
    assert len(bad) == 0, [(strs[2 * r + c], got[r, c], want[r, c]) for r, c in bad[:5]]


def test_device_parser_decimal_conversion_on_the_host(tmp_path):
    """csrc/kdejs_decimal.h is the conversion the DEVICE table parser uses; it compiles for the host too: three million
    spellings over its whole range against strtod, bit for bit, and the spellings it must leave to the host reader."""
    import shutil
    import subprocess

    from conftest import ROOT

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = tmp_path / "decimal_fuzz"
    subprocess.run([gxx, "-O2", "-std=c++17", os.path.join(ROOT, "tests", "decimal", "decimal_fuzz.cpp"), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe), "1000000", "7"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "mismatches 0" in out.stdout and "ACCEPTED" not in out.stdout


def test_file_bytes_reader_and_data_offset(nat, tmp_path):
    """kdejs_files_read (bytes of many files into slots, no parsing) and kdejs_tsv_data_offset (where the data rows of a
    table held as text begin: the header rule of kdejs_tsv_load)."""
    blobs = [b"core\tacc\n0.1\t0.2\n", b"0.5\t0.25\n" * 1000, b"", b"x" * 5000]
    paths = []
    for k, b in enumerate(blobs):
        q = tmp_path / f"f{k}.tsv"
        q.write_bytes(b)
        paths.append(str(q))
    arena = nat.TextArena(5, 4096 * 3, pinned=False)
    n = arena.read(paths + [str(tmp_path / "missing.tsv")], threads=3)
    assert n.tolist() == [len(blobs[0]), len(blobs[1]), 0, 5000, -1]
    for k in (0, 1, 3):
        assert arena.array[k, :n[k]].tobytes() == blobs[k]
    small = nat.TextArena(1, 100, pinned=False)
    assert small.read(paths[1:2]).tolist() == [-len(blobs[1])]
    L = nat.lib()

    def offset(text, skip):
        a = np.frombuffer(text, dtype=np.uint8)
        return L.kdejs_tsv_data_offset(a.ctypes.data, len(text), skip)

    assert offset(blobs[0], -1) == 9 and offset(blobs[0], 0) == 0 and offset(blobs[0], 1) == 9
    assert offset(b"0.1\t0.2\n0.3\t0.4\n", -1) == 0                     # numeric first row: nothing dropped
    assert offset(b"\n# note\nid\tcore\tacc\n1\t2\n", -1) == 20          # blank and comment lines are not the header
    assert offset(b"a\tb\nc\td", 2) == 7 and offset(b"a\tb", 5) == 3     # no final newline; more to skip than there is
    arena.close()
    small.close()


def test_number_vector_reader_matches_loadtxt(nat, tmp_path):
    """kdejs_numbers_load (the simulator's gene-frequency files, poppunk_mod.py:361) against np.loadtxt: a column, a
    tab-separated row, comments, a missing value; anything else is an error the caller can fall back from."""
    rng = np.random.default_rng(2)
    v = rng.random(3000)
    col = tmp_path / "col.txt"
    np.savetxt(col, v)
    row = tmp_path / "row.txt"
    row.write_text("# gene frequencies\n" + "\t".join(repr(float(x)) for x in v) + "\n")
    for p in (col, row):
        np.testing.assert_array_equal(nat.load_numbers(str(p)), np.loadtxt(p, delimiter="\t", dtype="float64").reshape(-1))
    mixed = tmp_path / "mixed.txt"
    mixed.write_text("0.5 1\tNA\r\n2e-3  # tail\n")
    got = nat.load_numbers(str(mixed))
    assert got[:2].tolist() == [0.5, 1.0] and np.isnan(got[2]) and got[3] == 2e-3 and got.shape == (4,)
    for bad, text in (("commas.txt", "1,2,3\n"), ("empty.txt", "\n# nothing\n")):
        q = tmp_path / bad
        q.write_text(text)
        with pytest.raises(ValueError):
            nat.load_numbers(str(q))
    from poppunk_mod_b200 import simulator

    np.testing.assert_array_equal(simulator._freqs(str(col)), v)
    commas = tmp_path / "c.txt"
    commas.write_text("0.25\n0.5\n")
    np.testing.assert_array_equal(simulator._freqs(str(commas)), [0.25, 0.5])
