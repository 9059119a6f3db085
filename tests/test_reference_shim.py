"""The unmodified reference (oracle/_ref, built from /root/reference on the host-memory OpenCL
shim) against the restated oracle.  On square grids the f32ref variant is the reference's own
arithmetic, so states and av_vels must agree BIT FOR BIT (only obstacle cells' rest population
differs by design: the reference zeroes it, quirk Q3, the canonical code keeps it)."""
import os

import numpy as np
import pytest

from oracle_bindings import REF_LIB, Oracle, Reference
from tools import cases

pytestmark = pytest.mark.skipif(not os.path.isfile(REF_LIB), reason="oracle/_ref not built")


def test_reference_timestep_equals_f32ref_oracle_bitwise(tmp_path):
    case = cases.shipped("128x128")
    pf, of = case.write(str(tmp_path))
    ref = Reference(pf, of, str(tmp_path))
    try:
        assert (ref.nx, ref.ny, ref.tot_cells) == (case.nx, case.ny, case.tot_cells)
        steps = 300
        av_ref = ref.steps(steps)
        f_ref = ref.download()
    finally:
        ref.close()
    o = Oracle("f32ref", case)
    f = o.init()
    av = o.run(f, steps).astype(np.float32)
    assert np.array_equal(av_ref, av)
    fluid = case.obstacles.ravel() == 0
    assert np.array_equal(f_ref[1:], f[1:])
    assert np.array_equal(f_ref[0][fluid], f[0][fluid])
    assert np.all(f_ref[0][~fluid] == 0.0)


def test_reference_from_perturbed_state(tmp_path):
    case = cases.random_case(64, 64, seed=3, walls=True)
    case.max_iters = 10
    pf, of = case.write(str(tmp_path))
    ref = Reference(pf, of, str(tmp_path))
    try:
        f0 = cases.perturbed_state(case, seed=3)
        ref.upload(f0)
        av_ref = ref.steps(25)
        f_ref = ref.download()
    finally:
        ref.close()
    o = Oracle("f32ref", case)
    f = f0.copy()
    av = o.run(f, 25).astype(np.float32)
    assert np.array_equal(av_ref, av)
    assert np.array_equal(f_ref[1:], f[1:])
