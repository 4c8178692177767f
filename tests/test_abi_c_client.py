"""The boundary is a C ABI: a plain-C program (gcc, no Python/torch/CUDA headers) links libkdejs.so."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "poppunk-mod_b200")
SRC = os.path.join(ROOT, "tests", "abi", "abi_client.c")


def _build(tmp_path):
    exe = str(tmp_path / "abi_client")
    subprocess.run(["/usr/bin/gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-L", PKG, "-lkdejs", f"-Wl,-rpath,{PKG}"], check=True)
    return exe


def _write_input(path, a, b):
    with open(path, "wb") as f:
        np.array([len(a), len(b)], dtype=np.int64).tofile(f)
        np.ascontiguousarray(a, dtype=np.float64).tofile(f)
        np.ascontiguousarray(b, dtype=np.float64).tofile(f)


def _no_gpu():
    import poppunk_mod_b200 as pk

    return pk._native.device_count() == 0


@pytest.mark.skipif(not _no_gpu(), reason="CPU-side check")
def test_c_client_links_and_fails_loudly_without_a_gpu(tmp_path):
    from oracle import workloads as wl

    exe = _build(tmp_path)
    a, b = wl.g1_inputs()
    _write_input(tmp_path / "in.bin", a[:50], b[:50])
    out = subprocess.run([exe, str(tmp_path / "in.bin"), "100", "0.25", "0.0", "0"], capture_output=True, text=True)
    assert out.returncode == 1                       # no device: KDEJS_ERR_CUDA, never a silent CPU answer
    assert "status 1" in out.stdout and "error" in out.stdout and "js" not in out.stdout.replace("kdejs", "")


@pytest.mark.gpu
def test_c_client_matches_the_oracle(tmp_path, golden):
    from oracle import workloads as wl

    exe = _build(tmp_path)
    a, b = wl.g1_inputs()
    _write_input(tmp_path / "in.bin", a, b)
    out = subprocess.run([exe, str(tmp_path / "in.bin"), "100", "0.25", "0.0", "0"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = dict(l.split(" ", 1) for l in out.stdout.strip().splitlines())
    want = golden["cases"]["G1_g025"]["oracle"]
    assert abs(float(lines["js"]) - want) <= 1e-10 * want
    rc, v0, v1 = lines["batch"].split()
    assert rc == "0" and float(v0) == float(lines["js"]) == float(v1)
