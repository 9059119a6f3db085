"""CPU-side checks of the product boundary: the C-ABI library loads and exports every symbol
include/lbm_b200.h declares, its pure-host entry points work, it refuses to run without a GPU
(no fallback), and the C host program keeps the reference's command line and diagnostics."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tools import cases


def test_library_exports_every_declared_symbol(lbm):
    lib = lbm.load()
    names = lbm.exported_symbols()
    assert len(names) >= 20 and "lbm_run" in names and "lbm_step" in names
    for n in names:
        assert hasattr(lib, n), n


def test_slab_rows_partition(lbm):
    for ny, world in [(16384, 8), (65536, 8), (128, 1), (130, 4), (7, 3)]:
        covered = []
        for r in range(world):
            y0, rows = lbm.slab_rows(ny, world, r)
            covered += list(range(y0, y0 + rows))
            assert abs(rows - ny / world) < 1
        assert covered == list(range(ny))
    with pytest.raises(lbm.LbmError):
        lbm.slab_rows(4, 8, 0)


def test_no_cpu_fallback(lbm):
    if lbm.device_count() > 0:
        pytest.skip("a GPU is present")
    c = cases.random_case(16, 16, seed=1)
    with pytest.raises(lbm.LbmError, match="no CUDA device"):
        lbm.Lattice(c.nx, c.ny, c.density, c.accel, c.omega, c.obstacles)


def run_exe(lbm, args, cwd):
    return subprocess.run([lbm.EXE_PATH] + args, cwd=cwd, capture_output=True, text=True)


def test_host_program_usage_and_diagnostics(lbm, tmp_path):
    # reference: usage() d2q9-bgk.c:941-945, die() :933-939, messages :490-528 and :606-624
    r = run_exe(lbm, [], tmp_path)
    assert r.returncode == 1 and r.stderr.startswith("Usage: ") and \
        r.stderr.rstrip().endswith("<paramfile> <obstaclefile>")
    r = run_exe(lbm, ["nope.params", "nope.dat"], tmp_path)
    assert r.returncode == 1 and "could not open input parameter file: nope.params" in r.stderr
    assert r.stderr.startswith("Error at line ")

    case = cases.random_case(8, 6, seed=2)
    pf, of = case.write(str(tmp_path))
    r = run_exe(lbm, [pf, "nope.dat"], tmp_path)
    assert "could not open input obstacles file: nope.dat" in r.stderr

    (tmp_path / "short.params").write_text("8\n6\n10\n")
    r = run_exe(lbm, ["short.params", of], tmp_path)
    assert r.returncode == 1 and "could not read param file: reynolds_dim" in r.stderr

    for text, msg in [("1 2\n", "expected 3 values per line in obstacle file"),
                      ("8 1 1\n", "obstacle x-coord out of range"),
                      ("1 6 1\n", "obstacle y-coord out of range"),
                      ("1 1 2\n", "obstacle blocked value should be 1")]:
        (tmp_path / "bad.dat").write_text(text)
        r = run_exe(lbm, [pf, "bad.dat"], tmp_path)
        assert r.returncode == 1 and msg in r.stderr, (text, r.stderr)


def test_host_program_fails_loudly_without_gpu(lbm, tmp_path):
    if lbm.device_count() > 0:
        pytest.skip("a GPU is present")
    case = cases.random_case(8, 6, seed=2)
    case.max_iters = 3
    pf, of = case.write(str(tmp_path))
    r = run_exe(lbm, [pf, of], tmp_path)
    assert r.returncode == 1 and "no CUDA device" in r.stderr
    assert not (tmp_path / "av_vels.dat").exists()


def test_parallel_obstacle_parser_equals_the_reference_loop(lbm, tmp_path):
    """the mmap + OpenMP parser of host/d2q9-bgk.c against the reference's fscanf loop
    (d2q9-bgk.c:615-628, kept behind LBM_SERIAL_PARSE=1): same obstacle map for well-formed files
    (with duplicates, blank lines, CRLF, no final newline) and, through the fallback, for files the
    fast scanner declines (three values spread over several lines)"""
    case = cases.random_case(300, 200, seed=4, fill=0.3)
    pf, of = case.write(str(tmp_path))
    body = open(of).read()
    variants = {
        "plain": body,
        "dups_blank_crlf": body + body[:4000] + "\n\n" + "5 7 1\r\n" + "9   11\t1  \n" + "12 13 1",
        "spread": "1 2\n1\n" + body,            # fscanf semantics: whitespace includes newlines
        "empty": "",
    }
    for name, text in variants.items():
        path = tmp_path / (name + ".dat")
        path.write_text(text)
        outs = []
        for serial in (False, True):
            env = dict(os.environ, LBM_PARSE_ONLY="1")
            if serial:
                env["LBM_SERIAL_PARSE"] = "1"
            r = subprocess.run([lbm.EXE_PATH, pf, str(path)], cwd=tmp_path, capture_output=True, text=True, env=env)
            assert r.returncode == 0, (name, r.stderr)
            outs.append(r.stdout)
        assert outs[0] == outs[1] and outs[0].startswith("parsed: 300 x 200"), (name, outs)
    # and errors keep the reference's diagnostics (first offending entry, found by the fallback)
    (tmp_path / "bad.dat").write_text(body + "300 1 1\n")
    r = subprocess.run([lbm.EXE_PATH, pf, "bad.dat"], cwd=tmp_path, capture_output=True, text=True,
                       env=dict(os.environ, LBM_PARSE_ONLY="1"))
    assert r.returncode == 1 and "obstacle x-coord out of range" in r.stderr
