"""tools/run_check.py: the native Python 3 restatement of the reference's check.py must agree
with the reference script executed unmodified (when the checkout is present)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tools import cases, run_check

REF_CHECKER = "/root/reference/check/check.py"
ROOT = cases.ROOT


def write_outputs(d, nx, ny, av, pressure):
    with open(os.path.join(d, "av_vels.dat"), "w") as f:
        for i, v in enumerate(av):
            f.write("%d:\t%.12E\n" % (i, v))
    with open(os.path.join(d, "final_state.dat"), "w") as f:
        for y in range(ny):
            for x in range(nx):
                f.write("%d %d %.12E %.12E %.12E %.12E %d\n" % (x, y, 0, 0, 0, pressure[y * nx + x], 0))


@pytest.mark.parametrize("scale,expect", [(1.0, 0), (1.004, 0), (1.02, 1)])
def test_native_check_verdicts(tmp_path, scale, expect):
    g = cases.golden("128x128")
    sim = tmp_path / "sim"
    sim.mkdir()
    write_outputs(str(sim), 128, 128, g["av_vels"][:] * scale, g["pressure"] * scale)
    npz = os.path.join(cases.GOLDEN_DIR, "128x128.npz")
    rc, d = run_check.native_check(npz, npz, str(sim / "av_vels.dat"), str(sim / "final_state.dat"),
                                   1.0, quiet=True)
    assert rc == expect
    if scale != 1.0:
        assert abs(abs(d["av_vels"]["max_diff_pcnt"]) - 100 * (scale - 1) / scale) < 1e-6


def test_nan_fails(tmp_path):
    g = cases.golden("128x128")
    av = g["av_vels"].copy()
    av[17] = np.nan
    write_outputs(str(tmp_path), 128, 128, av, g["pressure"])
    npz = os.path.join(cases.GOLDEN_DIR, "128x128.npz")
    rc, _ = run_check.native_check(npz, npz, str(tmp_path / "av_vels.dat"),
                                   str(tmp_path / "final_state.dat"), 1.0, quiet=True)
    assert rc == 1


@pytest.mark.skipif(not os.path.isfile(REF_CHECKER), reason="reference checkout not present")
@pytest.mark.parametrize("scale", [1.003, 1.03])
def test_native_agrees_with_unmodified_reference_checker(tmp_path, scale):
    g = cases.golden("128x128")
    write_outputs(str(tmp_path), 128, 128, g["av_vels"] * scale, g["pressure"] / scale)
    args = ["--ref-av-vels-file", "/root/reference/check/128x128.av_vels.dat",
            "--ref-final-state-file", "/root/reference/check/128x128.final_state.dat",
            "--av-vels-file", str(tmp_path / "av_vels.dat"),
            "--final-state-file", str(tmp_path / "final_state.dat")]
    tool = os.path.join(ROOT, "tools", "run_check.py")
    a = subprocess.run([sys.executable, tool] + args, capture_output=True, text=True)
    b = subprocess.run([sys.executable, tool, "--native"] + args, capture_output=True, text=True)
    assert a.returncode == b.returncode == (0 if scale < 1.01 else 1)
    assert "executing the reference checker" in a.stdout
    strip = lambda s: [l for l in s.splitlines() if not l.startswith("[run_check]")]
    assert strip(a.stdout) == strip(b.stdout)
