"""Pins the CPU oracle to the reference's golden vectors.

The f64 variant of oracle/canon must reproduce check/*.dat byte-for-byte: per-step here for the
first steps of every shipped case, full-length by sha256 for the small cases (all four with
LBM_FULL_GOLDEN=1; that was also done once when tests/golden/ was generated -- make_golden.py
asserts it).  The two float variants must stay inside check.py's 1 % of the golden.
"""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from oracle_bindings import CANON_EXE, Oracle
from tools import cases

MANIFEST = json.load(open(os.path.join(cases.GOLDEN_DIR, "manifest.json")))
REFERENCE = "/root/reference"


def sha256(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("name,steps", [("128x128", 1500), ("128x256", 800), ("256x256", 400),
                                        ("1024x1024", 40)])
def test_f64_oracle_reproduces_golden_av_vels_text(name, steps):
    case, gold = cases.shipped(name), cases.golden(name)
    o = Oracle("f64", case)
    av = o.run(o.init(), steps)
    # the golden file holds "%.12E" renderings; the oracle's doubles must render identically
    rendered = np.array([float("%.12E" % v) for v in av])
    assert np.array_equal(rendered, gold["av_vels"][:steps])


@pytest.mark.parametrize("name", ["128x128", "128x256"])
def test_f64_oracle_full_run_is_byte_identical(name, tmp_path):
    case = cases.shipped(name)
    pf, of = case.write(str(tmp_path))
    subprocess.check_call([CANON_EXE, "f64", pf, of, str(tmp_path)], stdout=subprocess.DEVNULL)
    assert sha256(tmp_path / "av_vels.dat") == MANIFEST[name]["av_vels"]["sha256"]
    assert sha256(tmp_path / "final_state.dat") == MANIFEST[name]["final_state"]["sha256"]


@pytest.mark.slow
@pytest.mark.skipif(not os.environ.get("LBM_FULL_GOLDEN"), reason="6 CPU-minutes; LBM_FULL_GOLDEN=1")
@pytest.mark.parametrize("name", ["256x256", "1024x1024"])
def test_f64_oracle_full_run_is_byte_identical_large(name, tmp_path):
    test_f64_oracle_full_run_is_byte_identical(name, tmp_path)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_manifest_matches_reference_files():
    for name, e in MANIFEST.items():
        av = "%s/check/%s.av_vels.dat" % (REFERENCE, name)
        assert sha256(av) == e["av_vels"]["sha256"]
        fs = "%s/check/%s.final_state.dat" % (REFERENCE, name)
        if os.path.isfile(fs):
            assert sha256(fs) == e["final_state"]["sha256"]
        else:
            assert "regenerated" in e["final_state"]["source"]
    # shipped inputs rebuilt from the fixtures are the reference's inputs
    for name in cases.SHIPPED:
        c = cases.shipped(name)
        vals = open("%s/input_%s.params" % (REFERENCE, name)).read().split()
        assert [c.nx, c.ny, c.max_iters, c.reynolds_dim] == [int(v) for v in vals[:4]]
        assert [c.density, c.accel, c.omega] == [float(v) for v in vals[4:]]
        ob = np.zeros((c.ny, c.nx), dtype=np.int32)
        for line in open("%s/obstacles_%s.dat" % (REFERENCE, name)):
            x, y, _ = (int(v) for v in line.split())
            ob[y, x] = 1
        assert np.array_equal(ob, c.obstacles)


def test_survey_pins_of_regenerated_goldens():
    # sha256 recorded at survey time for the two golden files missing from the checkout
    assert MANIFEST["256x256"]["final_state"]["sha256"] == \
        "5fef76c2744d48b3ce9deaf6bd216e5f53d73e9b83dd411250cf836095e2b592"
    assert MANIFEST["1024x1024"]["final_state"]["sha256"] == \
        "d89cd206dfd942954a001f50cac02b647425a739d5b3d45a8f1c733c78dfd8f9"


@pytest.mark.parametrize("variant", ["f32ref", "f32b200"])
@pytest.mark.parametrize("name,steps", [("128x128", 3000), ("128x256", 1500)])
def test_float_oracles_track_golden_within_check_tolerance(variant, name, steps):
    case, gold = cases.shipped(name), cases.golden(name)
    o = Oracle(variant, case)
    av = o.run(o.init(), steps).astype(np.float32).astype(np.float64)
    ref = gold["av_vels"][:steps]
    pct = 100.0 * (ref - av) / av
    assert np.all(np.isfinite(pct))
    assert np.max(np.abs(pct)) < 1.0           # check.py's tolerance (check/check.py:26-31)
    assert np.max(np.abs(pct)) < 0.05          # and in fact far inside it this early in the run


def test_full_length_f32b200_fixture_is_inside_tolerance():
    # av_vels of the full-length f32-strict oracle runs (stored in the fixtures) vs golden
    for name in cases.SHIPPED:
        g = cases.golden(name)
        sim = g["av_vels_f32b200"].astype(np.float64)
        pct = 100.0 * (g["av_vels"] - sim) / sim
        assert np.max(np.abs(pct)) < 0.2, name
