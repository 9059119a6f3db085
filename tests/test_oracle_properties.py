"""Properties of the CPU oracle that hold for any input (they guard the restatement itself, beyond
the golden pins): mass conservation, rest state is a fixed point without forcing, the three
arithmetic variants agree to rounding, mirror symmetry in y, and the acceleration mask."""
import numpy as np
import pytest

from oracle_bindings import Oracle
from tools import cases


@pytest.mark.parametrize("variant", ["f64", "f32ref", "f32b200"])
def test_mass_conserved_in_a_closed_box(variant):
    case = cases.random_case(48, 40, seed=2, walls=True)
    case.obstacles[:, 0] = case.obstacles[:, -1] = 1
    o = Oracle(variant, case)
    f = cases.perturbed_state(case, seed=2).astype(o.np_t)
    m0 = f.astype(np.float64).sum()
    o.run(f, 200)
    m1 = f.astype(np.float64).sum()
    assert abs(m1 - m0) / m0 < (1e-12 if variant == "f64" else 2e-5)


@pytest.mark.parametrize("variant", ["f64", "f32ref", "f32b200"])
def test_rest_state_without_forcing_is_a_fixed_point(variant):
    case = cases.random_case(32, 24, seed=4, accel=0.0)
    o = Oracle(variant, case)
    f = o.init()
    f0 = f.copy()
    av = o.run(f, 5)
    # equilibrium at rest relaxes onto itself up to one rounding of the weights
    assert np.max(np.abs(f - f0) / f0) < (1e-14 if variant == "f64" else 3e-7)   # 5 steps x 1 ulp
    assert np.all(av < 1e-6)


def test_variants_agree_to_rounding():
    case = cases.random_case(64, 32, seed=6, walls=True)
    f0 = cases.perturbed_state(case, seed=6)
    out = {}
    for v in ("f64", "f32ref", "f32b200"):
        o = Oracle(v, case)
        f = f0.astype(o.np_t)
        av = o.run(f, 50)
        out[v] = (f.astype(np.float64), av)
    for v in ("f32ref", "f32b200"):
        assert np.max(np.abs(out[v][0] - out["f64"][0]) / out["f64"][0]) < 5e-5
        assert np.max(np.abs(out[v][1] - out["f64"][1]) / out["f64"][1]) < 5e-5


def test_acceleration_only_touches_row_ny_minus_2_fluid_cells():
    case = cases.random_case(40, 16, seed=8)
    o = Oracle("f32b200", case)
    f = cases.perturbed_state(case, seed=8)
    g = f.copy()
    o._fn("accelerate")(case.nx, case.ny, o._real(case.density), o._real(case.accel), o._p(o.obst), o._p(g))
    changed = np.any(g != f, axis=0).reshape(case.ny, case.nx)
    assert not changed[: case.ny - 2].any() and not changed[case.ny - 1].any()
    assert not (changed[case.ny - 2] & (case.obstacles[case.ny - 2] != 0)).any()
    assert changed[case.ny - 2].any()
    # momentum goes in, mass does not
    assert abs(g.astype(np.float64).sum() - f.astype(np.float64).sum()) < 1e-6


def test_slab_and_channel_generators_are_deterministic(lbm):
    a = cases.channel(512, 256)
    b = cases.channel(512, 256)
    assert np.array_equal(a.obstacles, b.obstacles)
    assert a.obstacles[0].all() and a.obstacles[-1].all()
    frac = a.obstacles[1:-1].mean()
    assert 0.005 < frac < 0.03
    total = sum(cases.channel(512, 256, rows=lbm.slab_rows(256, 5, r)).shape[0] for r in range(5))
    assert total == 256
