"""GPU parity tests (run with -m gpu on a B200): the CUDA path, through the C ABI, against the
CPU oracle on the same seeded inputs.

Bars
  * state (all 9 populations of every cell): BIT-EXACT against the f32-strict oracle
    (oracle/canon_impl.h VARIANT_B200) -- integer comparison of the float bit patterns;
  * av_vels: the GPU sums cell speeds in double in a fixed tree, the oracle sequentially in
    double: relative difference <= 1e-12 on the un-narrowed values (reduction order only);
  * against the reference's own float arithmetic (f32ref oracle == unmodified reference, see
    test_reference_shim.py): relative difference <= 2e-4 on av_vels and on every population after
    200 steps (the restructured equilibrium expression rounds differently; measured 5.5e-5);
  * against the double-precision golden files: check.py's 1 % (tolerance stated in check.py:26-31).
"""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from oracle_bindings import Oracle
from tools import cases, run_check

pytestmark = pytest.mark.gpu
MANIFEST = json.load(open(os.path.join(cases.GOLDEN_DIR, "manifest.json")))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def make(lbm, case, **kw):
    return lbm.Lattice(case.nx, case.ny, case.density, case.accel, case.omega, case.obstacles, **kw)


def assert_state_bit_exact(f_gpu, f_cpu):
    same = bits(f_gpu) == bits(f_cpu)
    assert same.all(), "%d of %d values differ" % (np.count_nonzero(~same), same.size)


# sizes cover: 4-wide vectors with several rows per warp (nx < 128), one row per warp, many warps
# per row, 2-wide and scalar fallbacks (nx % 4 != 0), tiny and ragged grids, ny = 2 (the
# accelerated row is row 0 and both neighbours are the periodic image)
SIZES = [(128, 128), (64, 48), (256, 20), (1024, 12), (36, 50), (130, 37), (33, 17), (7, 5),
         (4, 3), (8, 2), (512, 3)]


@pytest.mark.parametrize("nx,ny", SIZES)
def test_single_steps_bit_exact(lbm, nx, ny):
    case = cases.random_case(nx, ny, seed=nx * 1000 + ny)
    f0 = cases.perturbed_state(case, seed=nx + ny)
    o = Oracle("f32b200", case)
    f = f0.copy()
    with make(lbm, case) as lat:
        lat.upload(f0)
        for _ in range(6):
            av_gpu = lat.step()
            f, av = o.step(f)
            assert_state_bit_exact(lat.download(), f)
            assert av_gpu == np.float32(av) or abs(float(av_gpu) - av) <= 1e-7 * abs(av)


@pytest.mark.parametrize("resident", ["0", "1"])
@pytest.mark.parametrize("nx,ny", SIZES)
def test_run_bit_exact_and_av_vels(lbm, nx, ny, resident, monkeypatch):
    # resident=0: launches; a short chunk so that graph replay, the direct-launch tail and the
    # un-accelerated last step are all exercised: 37 = 4 graph chunks of 8 + 5 direct.
    # resident=1: the persistent kernel (the default for lattices this small), 7 steps per launch so
    # that a run is several launches and the buffer parity flips between them: 37 = 5 x 7 + 2
    monkeypatch.setenv("LBM_CHUNK", "8")
    monkeypatch.setenv("LBM_RESIDENT", resident)
    monkeypatch.setenv("LBM_RES_CHUNK", "7")
    case = cases.random_case(nx, ny, seed=nx * 7 + ny, walls=(ny > 4))
    f0 = cases.perturbed_state(case, seed=ny)
    o = Oracle("f32b200", case)
    f = f0.copy()
    av = o.run(f, 37)
    with make(lbm, case) as lat:
        assert ("resident=smem" in lat.config) == (resident == "1")
        lat.upload(f0)
        av_gpu = lat.run(37, f64=True)
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu - av) / np.abs(av)) <= 1e-12
        # and a second run continues from the canonical state (no stale pre-acceleration)
        av2 = o.run(f, 9)
        av_gpu2 = lat.run(9, f64=True)
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu2 - av2) / np.abs(av2)) <= 1e-12


@pytest.mark.parametrize("tpb,vec", [(128, 4), (512, 4), (256, 2), (256, 1)])
def test_launch_variants_bit_exact(lbm, tpb, vec, monkeypatch):
    monkeypatch.setenv("LBM_TPB", str(tpb))
    monkeypatch.setenv("LBM_VEC", str(vec))
    monkeypatch.setenv("LBM_RESIDENT", "0")     # the one-step launches, not the persistent kernel
    case = cases.random_case(192, 40, seed=5, walls=True)
    f0 = cases.perturbed_state(case, seed=5)
    o = Oracle("f32b200", case)
    f = f0.copy()
    o.run(f, 20)
    with make(lbm, case) as lat:
        assert "vec=%d tpb=%d" % (vec, tpb) in lat.config
        lat.upload(f0)
        lat.run(20)
        assert_state_bit_exact(lat.download(), f)


def _resident_fits(nx, ny, rows, tpb):
    """mirror of decide_resident (lbm_engine.cu): shared memory of a block and cells per thread"""
    r = min(max(rows, -(-ny // 148)), ny)
    threads = min(min(512, max(32, tpb)), -(-r * nx // 32) * 32) // 32 * 32
    smem = 8 * (9 * r + 6) * nx + 64 * threads
    return smem <= 232448 and -(-r * nx // threads) <= 8, r


@pytest.mark.parametrize("tpb,rows", [(512, 1), (64, 1), (128, 3), (256, 2), (512, 5), (96, 7), (512, 40), (32, 2)])
@pytest.mark.parametrize("nx,ny", [(256, 96), (100, 31), (64, 20)])
def test_persistent_kernel_shapes_bit_exact(lbm, tpb, rows, nx, ny, monkeypatch):
    """lbm_resident_kernel: block shapes -- one row per block (every row is a boundary row), several
    rows with an interior, a ragged last block, 1 / 2 / 4 / 8 cells per thread, the whole lattice in ONE
    block (its own neighbour on both sides: 64 x 20 with 40 rows per block) -- one launch for the whole
    run, then a second run from the canonical state"""
    monkeypatch.setenv("LBM_RESIDENT", "1")
    monkeypatch.setenv("LBM_RES_ROWS", str(rows))
    monkeypatch.setenv("LBM_RES_TPB", str(tpb))
    case = cases.random_case(nx, ny, seed=nx + 3 * ny + rows, walls=True)
    f0 = cases.perturbed_state(case, seed=tpb)
    o = Oracle("f32b200", case)
    f = f0.copy()
    av = o.run(f, 60)
    fits, r = _resident_fits(nx, ny, rows, tpb)
    with make(lbm, case) as lat:
        if fits:
            assert "resident=smem(rows/block=%d," % r in lat.config, lat.config
        else:                              # too many rows for the shared memory of an SM, or > 8 cells per thread
            assert "resident=" not in lat.config
        lat.upload(f0)
        av_gpu = lat.run(60, f64=True)
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu - av) / np.abs(av)) <= 1e-12
        av2 = o.run(f, 3)
        av_gpu2 = lat.run(3, f64=True)
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu2 - av2) / np.abs(av2)) <= 1e-12


def test_persistent_kernel_is_the_default_for_the_small_shipped_cases(lbm):
    for name, want in (("128x128", True), ("128x256", True), ("256x256", True), ("1024x1024", False)):   # 1024^2: too big for shared memory
        with make(lbm, cases.shipped(name)) as lat:
            assert ("resident=smem" in lat.config) == want, lat.config


def test_persistent_kernel_equals_launches_on_a_long_run(lbm, monkeypatch):
    """5000 steps of the shipped 256x256 case: the persistent kernel (5 launches) and the one-step
    launches give the same bits -- thousands of neighbour handshakes without a single stale read"""
    case = cases.shipped("256x256")
    out = {}
    for resident in ("1", "0"):
        monkeypatch.setenv("LBM_RESIDENT", resident)
        with make(lbm, case) as lat:
            lat.init_equilibrium()
            out[resident] = (lat.run(5000, f64=True), lat.download())
    assert_state_bit_exact(out["1"][1], out["0"][1])
    assert np.max(np.abs(out["1"][0] - out["0"][0]) / out["0"][0]) <= 1e-12


def test_persistent_kernel_falls_back_to_launches_when_it_cannot_be_placed(lbm, monkeypatch):
    """a cooperative launch the device refuses (simulated: e.g. an MPS partition smaller than the
    occupancy query assumed) is not an error: nothing ran, the handle carries on with one launch per step"""
    monkeypatch.setenv("LBM_TEST_RESIDENT_LAUNCH_FAIL", "1")
    case = cases.random_case(128, 64, seed=11, walls=True)
    f0 = cases.perturbed_state(case, seed=11)
    o = Oracle("f32b200", case)
    f = f0.copy()
    av = o.run(f, 30)
    with make(lbm, case) as lat:
        assert "resident=smem" in lat.config
        lat.upload(f0)
        av_gpu = lat.run(30, f64=True)
        assert "resident=" not in lat.config
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu - av) / np.abs(av)) <= 1e-12


def test_persistent_kernel_times_out_instead_of_hanging(lbm, monkeypatch):
    """a block that never publishes its progress: its neighbours give up after ~1 s of SM clocks, the
    launch ends, lbm_run reports the failure and the handle refuses further runs"""
    monkeypatch.setenv("LBM_RESIDENT", "1")
    monkeypatch.setenv("LBM_TEST_RESIDENT_STALL", "3")
    case = cases.random_case(128, 64, seed=9, walls=True)
    with make(lbm, case) as lat:
        lat.init_equilibrium()
        with pytest.raises(RuntimeError, match="timed out"):
            lat.run(10)
        with pytest.raises(RuntimeError, match="earlier run"):
            lat.run(2)


def test_device_init_equals_host_init(lbm):
    case = cases.shipped("128x256")
    with make(lbm, case) as lat:
        lat.init_equilibrium()
        assert_state_bit_exact(lat.download(), case.initial_state())


def test_against_reference_float_arithmetic(lbm):
    case = cases.shipped("128x128")
    o = Oracle("f32ref", case)
    f = o.init()
    av_ref = o.run(f, 200).astype(np.float32).astype(np.float64)
    with make(lbm, case) as lat:
        lat.init_equilibrium()
        av = lat.run(200).astype(np.float64)
        f_gpu = lat.download()
    assert np.max(np.abs(av - av_ref) / av_ref) <= 2e-4
    fluid = case.obstacles.ravel() == 0
    assert np.max(np.abs(f_gpu[:, fluid] - f[:, fluid]) / f[:, fluid]) <= 2e-4


def test_av_velocity_and_macroscopic(lbm):
    case = cases.shipped("128x128")
    o = Oracle("f32b200", case)
    f = o.init()
    o.run(f, 300)
    with make(lbm, case) as lat:
        lat.init_equilibrium()
        lat.run(300)
        # the reference's av_velocity accumulates in float (d2q9-bgk.c:429,466), the GPU in double
        assert abs(float(lat.av_velocity()) - o.av_velocity(f)) <= 1e-5 * o.av_velocity(f)
        m = lat.macroscopic()
    # write_values' float expressions (d2q9-bgk.c:857-897) are mirrored exactly
    assert_state_bit_exact(m, o.macroscopic(f))


def test_mass_is_conserved_and_obstacle_nan_hazard(lbm):
    # closed box: total mass changes only by float rounding; a zero-density obstacle cell must
    # not leak NaNs (the reference's 0/1-multiply select would: kernels.cl:179-198)
    case = cases.shipped("128x128")
    f0 = case.initial_state()
    ob = case.obstacles.ravel() != 0
    f0[:, np.flatnonzero(ob)[:50]] = 0.0
    with make(lbm, case) as lat:
        lat.upload(f0)
        av = lat.run(500)
        f = lat.download()
        mass_gpu = lat.total_density()
    assert np.all(np.isfinite(av)) and np.all(np.isfinite(f))
    m0, m1 = f0.astype(np.float64).sum(), f.astype(np.float64).sum()
    assert abs(m1 - m0) / m0 < 1e-5
    assert abs(mass_gpu - m1) <= 1e-12 * m1          # total_density (d2q9-bgk.c:822-838) on the device


@pytest.mark.parametrize("name", cases.SHIPPED)
def test_shipped_case_full_length_through_the_executable(lbm, name, tmp_path):
    """d2q9-bgk.exe <paramfile> <obstaclefile> on every shipped case, full iteration count:
    outputs pass the checker against the golden fixture, and are byte-identical to what the
    f32-strict oracle wrote for the same run (sha256 in tests/golden/manifest.json)."""
    case = cases.shipped(name)
    pf, of = case.write(str(tmp_path))
    r = subprocess.run([lbm.EXE_PATH, pf, of], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "==done==" and lines[1].startswith("Reynolds number:\t\t")
    assert lines[2].startswith("Elapsed time:\t\t\t") and lines[3].startswith("Elapsed user CPU time:\t\t")
    assert lines[4].startswith("Elapsed system CPU time:\t")
    npz = os.path.join(cases.GOLDEN_DIR, name + ".npz")
    rc, d = run_check.native_check(npz, npz, str(tmp_path / "av_vels.dat"),
                                   str(tmp_path / "final_state.dat"), 1.0, quiet=True)
    assert rc == 0, d
    sha = lambda p: hashlib.sha256(open(p, "rb").read()).hexdigest()
    assert sha(tmp_path / "final_state.dat") == MANIFEST[name]["f32b200"]["final_state_sha256"]
    # av_vels: identical up to the reduction order of the per-step double sum
    av = np.loadtxt(tmp_path / "av_vels.dat", usecols=[1])
    ref = cases.golden(name)["av_vels_f32b200"].astype(np.float64)
    assert np.max(np.abs(av - ref) / ref) <= 1.2e-7
    print("%s: av_vels worst %.3g %%, pressure worst %.3g %%, av_vels.dat byte-identical to oracle: %s"
          % (name, d["av_vels"]["max_diff_pcnt"], d["final_state"]["max_diff_pcnt"],
             sha(tmp_path / "av_vels.dat") == MANIFEST[name]["f32b200"]["av_vels_sha256"]))


def test_step_loop_equals_run_at_1024(lbm):
    case = cases.shipped("1024x1024")
    with make(lbm, case) as a, make(lbm, case) as b:
        a.init_equilibrium()
        b.init_equilibrium()
        av_a = a.run(24)
        av_b = np.array([b.step() for _ in range(24)], dtype=np.float32)
        assert np.array_equal(av_a, av_b)
        assert_state_bit_exact(a.download(), b.download())


def test_full_size_channel_properties(lbm):
    """BASELINE.json's 16384x16384 synthetic channel: properties that need no oracle at this
    size -- mass conservation, finite positive averages, flow accelerates from rest, and
    determinism (two engines, bit-identical states and averages)."""
    nx = ny = 16384
    ob = cases.channel(nx, ny, rows=(0, ny))
    lat = lbm.Lattice(nx, ny, 0.1, 0.005, 1.85, ob)
    try:
        assert "fuse=2" in lat.config and "stream=tma" in lat.config   # HBM-streaming slab: two timesteps per pass
        lat.init_equilibrium()
        av = lat.run(12, f64=True)
        m = lat.macroscopic()
        fluid = ob.ravel() == 0
        mass = (3.0 * m[3].astype(np.float64))[fluid].sum()      # pressure = rho / 3
        expect = np.float64(np.float32(0.1)) * fluid.sum()
        assert abs(mass - expect) / expect < 1e-5
        assert np.all(np.isfinite(av)) and np.all(av > 0) and np.all(np.diff(av) > 0)
        lat.init_equilibrium()
        av2 = lat.run(12, f64=True)
        assert np.array_equal(av, av2)
        # crop parity: rows far from the obstacles-free walls are not needed -- compare the whole
        # first 64 rows of a 16384x64 channel with the oracle instead (same nx, same kernel path)
    finally:
        lat.close()
    small = cases.channel(nx, 64)
    o = Oracle("f32b200", small)
    f = o.init()
    avo = o.run(f, 11)
    for fuse in ("1", "2"):           # the one-step kernel and the two-step passes the big grid uses
        os.environ["LBM_FUSE"] = fuse
        try:
            with make(lbm, small) as lat2:
                assert ("fuse=%s" % fuse) in lat2.config
                lat2.init_equilibrium()
                avg = lat2.run(11, f64=True)
                assert_state_bit_exact(lat2.download(), f)
        finally:
            del os.environ["LBM_FUSE"]
        assert np.max(np.abs(avg - avo) / avo) <= 1e-12


# ---- more than one GPU (skipped on a single-GPU box) -------------------------------------------
def _gpus(lbm):
    return lbm.device_count()


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("ngpus", [2, 4, 8])
def test_one_process_row_slabs_bit_exact(lbm, ngpus):
    """lbm_create(ngpus=N): slabs on N GPUs of this process, halos by NVLink peer stores"""
    if _gpus(lbm) < ngpus:
        pytest.skip("needs %d GPUs" % ngpus)
    case = cases.random_case(256, 100, seed=21, walls=True)      # 100 rows: ragged split for N=8
    f0 = cases.perturbed_state(case, seed=21)
    o = Oracle("f32b200", case)
    f = f0.copy()
    av = o.run(f, 30)
    with make(lbm, case, ngpus=ngpus) as lat:
        assert "slabs=%d" % ngpus in lat.config
        lat.upload(f0)
        av_gpu = lat.run(30, f64=True)
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu - av) / np.abs(av)) <= 1e-12
        av2 = o.run(f, 7)
        av_gpu2 = np.array([lat.step() for _ in range(7)])
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu2 - av2) / av2) <= 1e-6
        assert_state_bit_exact(lat.macroscopic(), o.macroscopic(f))


@pytest.mark.parametrize("world,halo,nx", [(2, "p2p", 512), (2, "p2p-kernels", 512), (2, "p2p", 516),
                                           (2, "nccl", 512), (2, "p2p-allreduce", 512),
                                           (3, "p2p", 1024), (4, "p2p", 512),
                                           (8, "p2p", 512), (8, "p2p-kernels", 512)])
def test_one_process_per_gpu_bit_exact(lbm, world, halo, nx):
    """lbm_create_rank under torch.distributed.run.  p2p: halos by peer stores into CUDA-IPC mapped
    neighbour memory, ring ordering by the step kernel's own boundary blocks (nx % 32 == 0) or by
    wait/signal kernels (nx = 516, or LBM_RING=kernels); nccl: ncclSend/ncclRecv halos."""
    if _gpus(lbm) < world:
        pytest.skip("needs %d GPUs" % world)
    import sys
    env = dict(os.environ, LBM_HALO="nccl" if halo == "nccl" else "p2p")
    if halo == "p2p-kernels":
        env["LBM_RING"] = "kernels"
    if halo == "p2p-allreduce":
        env["LBM_REDUCE"] = "step"      # per-step allreduce of the speed sum (north-star wording), done in-kernel
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(cases.ROOT, "tools", "multirank_check.py"), "--nx", str(nx), "--ny", "100",
           "--steps", "40"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    print(r.stdout[-600:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "-> OK" in r.stdout
    want = {"nccl": "nccl-sendrecv", "p2p-kernels": "wait/signal-kernels",
            "p2p-allreduce": "reduce=in-kernel-allreduce-per-step",
            "p2p": "in-kernel-ring" if nx % 32 == 0 else "wait/signal-kernels"}[halo]
    assert want in r.stdout


# ---- edge cases of the acceleration mask and of the geometry ------------------------------------
@pytest.mark.parametrize("accel,blocked_row", [(1.0, False), (3.0, False), (0.01, True)])
def test_acceleration_mask_edge_cases(lbm, accel, blocked_row):
    """accel = 1.0 puts w1 = density*accel/9 right at the size of f3, so the `f - w > 0` test of
    kernels.cl:29-32 is true for some cells and false for others; accel = 3.0 makes it false
    everywhere; a fully blocked row ny-2 must leave the flow unforced"""
    case = cases.random_case(96, 24, seed=31, accel=accel)
    if blocked_row:
        case.obstacles[case.ny - 2, :] = 1
    f0 = cases.perturbed_state(case, seed=31, amp=0.2)
    o = Oracle("f32b200", case)
    f = f0.copy()
    av = o.run(f, 25)
    with make(lbm, case) as lat:
        lat.upload(f0)
        av_gpu = lat.run(25, f64=True)
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu - av) / np.abs(av)) <= 1e-12


@pytest.mark.parametrize("nx,ny", [(4, 64), (8, 33), (1000, 37), (2, 9), (1, 6), (3, 3)])
def test_narrow_and_ragged_grids(lbm, nx, ny):
    case = cases.random_case(nx, ny, seed=nx * 31 + ny, fill=0.05)
    f0 = cases.perturbed_state(case, seed=nx)
    if case.tot_cells == 0:
        pytest.skip("all cells blocked")
    o = Oracle("f32b200", case)
    f = f0.copy()
    o.run(f, 11)
    with make(lbm, case) as lat:
        lat.upload(f0)
        lat.run(11)
        assert_state_bit_exact(lat.download(), f)


def test_zero_iterations_and_reupload(lbm):
    case = cases.random_case(64, 16, seed=3)
    f0 = cases.perturbed_state(case, seed=3)
    with make(lbm, case) as lat:
        lat.upload(f0)
        assert lat.run(0).size == 0
        assert_state_bit_exact(lat.download(), f0)
        lat.run(3)
        lat.upload(f0)                         # a fresh upload resets the ghost rows too
        o = Oracle("f32b200", case)
        f = f0.copy()
        o.run(f, 4)
        lat.run(4)
        assert_state_bit_exact(lat.download(), f)


def test_executable_on_two_gpus(lbm, tmp_path):
    """LBM_GPUS=2: same outputs as one GPU, byte for byte (state) / to reduction order (av_vels)"""
    if _gpus(lbm) < 2:
        pytest.skip("needs 2 GPUs")
    case = cases.shipped("128x256")
    outs = {}
    for n in (1, 2):
        d = tmp_path / ("g%d" % n)
        d.mkdir()
        pf, of = case.write(str(d), iters=3000)
        r = subprocess.run([lbm.EXE_PATH, pf, of], cwd=d, capture_output=True, text=True,
                           env=dict(os.environ, LBM_GPUS=str(n)))
        assert r.returncode == 0, r.stderr
        outs[n] = (open(d / "final_state.dat", "rb").read(), np.loadtxt(d / "av_vels.dat", usecols=[1]))
    assert outs[1][0] == outs[2][0]
    assert np.max(np.abs(outs[1][1] - outs[2][1]) / outs[1][1]) <= 1.2e-7


# ---- S timesteps per pass: the TMA / mbarrier streaming kernel (LBM_FUSE=2|3|4) --------------------
# widths: one strip, several strips, a last strip of 8 / 16 / 64 columns (16384 = 136 * 120 + 64), widths
# that are a multiple of 16 but not of 32 or 120; heights: one short tile, ragged last tile, ny = 16
# (the minimum: 4 ghost rows on either side of every slab are recomputed from the neighbour's rows)
STREAM_SIZES = [(128, 128), (256, 20), (1024, 16), (512, 37), (144, 33), (1008, 33), (240, 16), (2048, 70),
                (16384, 19)]


def _run_and_compare(lbm, case, f0, iters, more=5):
    o = Oracle("f32b200", case)
    f = f0.copy()
    av = o.run(f, iters)
    with make(lbm, case) as lat:
        lat.upload(f0)
        av_gpu = lat.run(iters, f64=True)
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu - av) / np.abs(av)) <= 1e-12
        av2 = o.run(f, more)
        av_gpu2 = lat.run(more, f64=True)       # continue from the canonical state: passes + plain steps
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu2 - av2) / np.abs(av2)) <= 1e-12
        return lat.config


@pytest.mark.parametrize("nx,ny", STREAM_SIZES)
@pytest.mark.parametrize("iters", [2, 7, 40])
def test_two_step_passes_bit_exact(lbm, nx, ny, iters, monkeypatch):
    """temporal blocking through the TMA-fed shared-memory pipeline: same arithmetic per cell and step,
    so still bit-identical to the oracle; odd counts end with one ordinary step"""
    monkeypatch.setenv("LBM_FUSE", "2")
    monkeypatch.setenv("LBM_CHUNK", "12")
    case = cases.random_case(nx, ny, seed=nx + 3 * ny, walls=(ny > 8))
    f0 = cases.perturbed_state(case, seed=nx)
    cfg = _run_and_compare(lbm, case, f0, iters)
    assert "fuse=2" in cfg and "stream=tma(S=2" in cfg


@pytest.mark.parametrize("fuse,cfg,tile_h", [(3, None, None), (4, None, None), (2, 2, None), (2, 4, None),
                                             (3, 6, None), (2, 0, 10), (3, 1, 8), (2, 5, None)])
@pytest.mark.parametrize("nx,ny", [(512, 37), (1008, 64), (16384, 19)])
def test_stream_kernel_shapes_bit_exact(lbm, fuse, cfg, tile_h, nx, ny, monkeypatch):
    """every instantiated shape of the streaming kernel (steps per pass, warps per group, TMA stages),
    small tile heights (many tiles per strip: the overlap rows between tiles are recomputed)"""
    monkeypatch.setenv("LBM_FUSE", str(fuse))
    monkeypatch.setenv("LBM_CHUNK", "12")
    if cfg is not None:
        monkeypatch.setenv("LBM_STREAM_CFG", str(cfg))
    if tile_h is not None:
        monkeypatch.setenv("LBM_TILE_H", str(tile_h))
    case = cases.random_case(nx, ny, seed=nx + 5 * ny + fuse, walls=True)
    f0 = cases.perturbed_state(case, seed=ny)
    got = _run_and_compare(lbm, case, f0, 25, more=7)
    assert "stream=tma(S=" in got


@pytest.mark.parametrize("nx,ny", [(136, 40), (1000, 33), (248, 64), (512, 12), (64, 64)])
def test_streaming_falls_back_to_the_one_step_kernel(lbm, nx, ny, monkeypatch):
    """widths that are not a multiple of 16 (TMA row pitch), nx < 128, or fewer than 16 rows per slab"""
    monkeypatch.setenv("LBM_FUSE", "2")
    case = cases.random_case(nx, ny, seed=nx + ny, walls=True)
    f0 = cases.perturbed_state(case, seed=nx)
    assert "fuse=1" in _run_and_compare(lbm, case, f0, 9)


def test_two_step_passes_shipped_case(lbm, monkeypatch):
    monkeypatch.setenv("LBM_FUSE", "2")
    case = cases.shipped("128x256")             # periodic in y, accelerated row next to the seam
    o = Oracle("f32b200", case)
    f = o.init()
    av = o.run(f, 501)
    with make(lbm, case) as lat:
        lat.init_equilibrium()
        av_gpu = lat.run(501, f64=True)
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu - av) / np.abs(av)) <= 1e-12


def test_periodic_seam_without_walls(lbm, monkeypatch):
    """no walls at y = 0 / ny-1: the flow crosses the periodic seam, so the ghost zones (depth 4, all nine
    planes, obstacle and acceleration flags of the mirrored rows) carry real data every pass"""
    for fuse in ("2", "3", "4"):
        monkeypatch.setenv("LBM_FUSE", fuse)
        case = cases.random_case(256, 48, seed=77, walls=False, fill=0.03)
        f0 = cases.perturbed_state(case, seed=78, amp=0.1)
        assert ("fuse=%s" % fuse) in _run_and_compare(lbm, case, f0, 36)


def _torchrun(world, script_args, env, timeout=300):
    import sys
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(cases.ROOT, "tools", "multirank_check.py")] + script_args
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=timeout)


@pytest.mark.parametrize("world,fuse,reduce,tile_h", [
    (2, 2, "batched", None), (3, 2, "batched", None), (8, 2, "batched", None), (2, 3, "batched", None),
    (4, 4, "batched", None), (8, 3, "step", None), (2, 2, "step", None), (4, 2, "step", None),
    # tiles shorter than the ghost depth: several rows of tiles touch a slab's ghost zones and all of
    # them take part in the ring ordering
    (2, 2, "batched", 2), (2, 3, "step", 2), (2, 4, "batched", 2), (8, 3, "batched", 2)])
def test_stream_passes_one_process_per_gpu(lbm, world, fuse, reduce, tile_h):
    if _gpus(lbm) < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, LBM_FUSE=str(fuse), LBM_CHUNK="12")
    if tile_h is not None:
        env["LBM_TILE_H"] = str(tile_h)
    if reduce == "step":
        env["LBM_REDUCE"] = "step"
    r = _torchrun(world, ["--nx", "512", "--ny", "%d" % (50 * world + 3), "--steps", "41"], env)
    print(r.stdout[-600:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "-> OK" in r.stdout and ("fuse=%d" % fuse) in r.stdout
    if reduce == "step":
        assert "in-kernel-allreduce-per-step" in r.stdout


@pytest.mark.parametrize("ngpus,fuse", [(2, 2), (4, 3), (8, 2)])
def test_stream_passes_one_process_row_slabs(lbm, ngpus, fuse, monkeypatch):
    """lbm_create(ngpus=N) -- the mode d2q9-bgk.exe uses with LBM_GPUS=N -- with the streaming kernel:
    same in-kernel ring as the one-process-per-GPU mode, peer pointers by cudaDeviceEnablePeerAccess"""
    if _gpus(lbm) < ngpus:
        pytest.skip("needs %d GPUs" % ngpus)
    monkeypatch.setenv("LBM_FUSE", str(fuse))
    monkeypatch.setenv("LBM_CHUNK", "12")
    case = cases.random_case(512, 40 * ngpus + 5, seed=21, walls=True)
    f0 = cases.perturbed_state(case, seed=21)
    o = Oracle("f32b200", case)
    f = f0.copy()
    av = o.run(f, 31)
    with make(lbm, case, ngpus=ngpus) as lat:
        assert "slabs=%d" % ngpus in lat.config and ("fuse=%d" % fuse) in lat.config
        lat.upload(f0)
        av_gpu = lat.run(31, f64=True)
        assert_state_bit_exact(lat.download(), f)
        assert np.max(np.abs(av_gpu - av) / np.abs(av)) <= 1e-12


def test_ring_timeout_is_an_error_not_a_hang(lbm):
    """negative test: rank 1 never publishes its progress (LBM_TEST_RING_STALL=1); its neighbours give up
    after the bounded spin and every rank's lbm_run returns the time-out error"""
    if _gpus(lbm) < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, LBM_FUSE="2", LBM_TEST_RING_STALL="1")
    r = _torchrun(2, ["--nx", "512", "--ny", "100", "--steps", "8", "--expect-timeout"], env, timeout=240)
    print(r.stdout[-600:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "timed out waiting for a neighbour GPU" in r.stdout and "-> OK" in r.stdout
