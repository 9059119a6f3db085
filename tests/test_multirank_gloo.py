"""World-size-2/3 CPU model (gloo) of the multi-GPU protocol: the same slab arithmetic
(lbm_slab_rows from the C ABI), the same ring-periodic halo (rows' populations 4,7,8 travel down,
2,5,6 travel up), the accelerated row living on whichever rank owns global row ny-2, and the
per-step speed totals combined in rank order.  Each rank advances its padded slab with the
f32-strict oracle; the assembled lattice must equal the oracle run on the whole grid bit for bit.
This pins the host-side decomposition logic that bench.py / tools/multirank_check.py and the
engine's lbm_create_rank share, without a GPU.  A second model does the same for the protocol of the
streaming kernel (S timesteps per pass: ghost depth 4 with all nine planes, the ghost zones recomputed
redundantly inside a pass, one store of the boundary rows into the neighbours per pass)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LO_PLANES, HI_PLANES = (4, 7, 8), (2, 5, 6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _accelerate_row(case, f3d, row, obst_row):
    """numpy restatement of kernels.cl:17-41 on one row of a [9, rows, nx] float32 array"""
    d, a = np.float32(case.density), np.float32(case.accel)
    a1 = np.float32(np.float64(d * a) / 9.0)
    a2 = np.float32(np.float64(d * a) / 36.0)
    r = f3d[:, row, :]
    m = (obst_row == 0) & ((r[3] - a1) > 0) & ((r[6] - a2) > 0) & ((r[7] - a2) > 0)
    r[1][m] += a1; r[5][m] += a2; r[8][m] += a2
    r[3][m] -= a1; r[6][m] -= a2; r[7][m] -= a2


def _worker(rank, world, port, nx, ny, steps, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(1, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle_bindings import Oracle
    from tools import cases
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lbm = importlib.import_module("hpc-lattice-boltzmann_b200")

    case = cases.random_case(nx, ny, seed=5, walls=True)
    f0 = cases.perturbed_state(case, seed=5).reshape(9, ny, nx)
    y0, rows = lbm.slab_rows(ny, world, rank)
    lo, hi = (rank - 1) % world, (rank + 1) % world

    # padded slab: ghost row below (0) and above (rows+1); a sub-lattice the oracle can step
    pad_ob = np.zeros((rows + 2, nx), dtype=np.int32)
    pad_ob[1:-1] = case.obstacles[y0:y0 + rows]
    sub = cases.Case("slab", nx, rows + 2, 0, 10, case.density, case.accel, case.omega, pad_ob)
    o = Oracle("f32b200", sub)
    f = np.zeros((9, rows + 2, nx), dtype=np.float32)
    f[:, 1:-1] = f0[:, y0:y0 + rows]
    totals = np.zeros(steps)

    def exchange(buf):
        down = torch.from_numpy(np.ascontiguousarray(buf[list(LO_PLANES), 1]))
        up = torch.from_numpy(np.ascontiguousarray(buf[list(HI_PLANES), rows]))
        got_hi, got_lo = torch.empty_like(down), torch.empty_like(up)
        reqs = [dist.isend(down, lo, tag=1), dist.isend(up, hi, tag=2),
                dist.irecv(got_hi, hi, tag=1), dist.irecv(got_lo, lo, tag=2)]
        for r in reqs:
            r.wait()
        buf[list(LO_PLANES), rows + 1] = got_hi.numpy()     # my upper ghost <- upper rank's first row
        buf[list(HI_PLANES), 0] = got_lo.numpy()            # my lower ghost <- lower rank's last row

    for t in range(steps):
        if y0 <= ny - 2 < y0 + rows:
            _accelerate_row(case, f, ny - 2 - y0 + 1, pad_ob[ny - 2 - y0 + 1])
        exchange(f)
        nxt, _ = o.step(np.ascontiguousarray(f.reshape(9, -1)), accel=False)
        sp = o._speeds.reshape(rows + 2, nx)[1:-1]
        totals[t] = sp.astype(np.float64).sum()
        f = nxt.reshape(9, rows + 2, nx)

    gathered = [torch.zeros(steps, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(totals))
    av = sum(g.numpy() for g in gathered) / case.tot_cells        # rank order
    np.save(os.path.join(out_dir, "slab%d.npy" % rank), f[:, 1:-1])
    if rank == 0:
        np.save(os.path.join(out_dir, "av.npy"), av)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny", [(2, 48, 40), (3, 32, 31)])
def test_row_slab_protocol_matches_whole_grid_oracle(world, nx, ny, tmp_path, lbm):
    import torch.multiprocessing as mp
    from oracle_bindings import Oracle
    from tools import cases
    steps = 25
    mp.spawn(_worker, args=(world, _free_port(), nx, ny, steps, str(tmp_path)), nprocs=world, join=True)
    case = cases.random_case(nx, ny, seed=5, walls=True)
    f = cases.perturbed_state(case, seed=5)
    o = Oracle("f32b200", case)
    av = o.run(f, steps)
    got = np.concatenate([np.load(tmp_path / ("slab%d.npy" % r)) for r in range(world)], axis=1)
    assert np.array_equal(got.reshape(9, -1).view(np.uint32), f.view(np.uint32))
    assert np.max(np.abs(np.load(tmp_path / "av.npy") - av) / av) <= 1e-12


def test_channel_slab_generation_matches_global_map(lbm):
    from tools import cases
    full = cases.channel(256, 192)
    for world in (1, 2, 3, 8):
        for r in range(world):
            y0, rows = lbm.slab_rows(192, world, r)
            assert np.array_equal(cases.channel(256, 192, rows=(y0, rows)), full.obstacles[y0:y0 + rows])


# ---- the S-timesteps-per-pass protocol (what lbm_stream_kernel does across ranks) ---------------
GHOST = 4


def _worker_stream(rank, world, port, nx, ny, S, passes, out_dir):
    """CPU model of one rank of the streaming kernel's multi-GPU protocol: the slab carries GHOST ghost
    rows on either side with ALL nine planes and the obstacle / acceleration flags of the lattice rows
    they mirror; a pass advances the whole padded slab S timesteps (rows within S of the padded edge
    go stale -- they are never used), sums speeds over the owned rows only, applies the following
    step's acceleration to every copy of lattice row ny-2, and afterwards each rank stores its first /
    last GHOST owned rows into its neighbours' ghost zones.  No exchange inside a pass."""
    sys.path.insert(0, ROOT)
    sys.path.insert(1, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle_bindings import Oracle
    from tools import cases
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lbm = importlib.import_module("hpc-lattice-boltzmann_b200")

    case = cases.random_case(nx, ny, seed=9, walls=False, fill=0.05)     # flow crosses the periodic seam
    f0 = cases.perturbed_state(case, seed=9).reshape(9, ny, nx)
    y0, rows = lbm.slab_rows(ny, world, rank)
    lo, hi = (rank - 1) % world, (rank + 1) % world
    G = GHOST
    gy = (np.arange(y0 - G, y0 + rows + G) % ny)                         # lattice row of every padded row
    pad_ob = case.obstacles[gy]
    slab = Oracle("f32b200", cases.Case("slab", nx, rows + 2 * G, 0, 10, case.density, case.accel, case.omega, pad_ob))
    accel_rows = [r for r in range(rows + 2 * G) if gy[r] == ny - 2]

    def accelerate(f3d):
        for r in accel_rows:
            _accelerate_row(case, f3d, r, pad_ob[r])

    def refresh(f3d):
        """my first / last G owned rows -> the neighbours' ghost zones (all planes)"""
        down = torch.from_numpy(np.ascontiguousarray(f3d[:, G:2 * G]))
        up = torch.from_numpy(np.ascontiguousarray(f3d[:, rows:rows + G]))
        from_hi, from_lo = torch.empty_like(down), torch.empty_like(up)
        reqs = [dist.isend(down, lo, tag=1), dist.isend(up, hi, tag=2),
                dist.irecv(from_hi, hi, tag=1), dist.irecv(from_lo, lo, tag=2)]
        for r in reqs:
            r.wait()
        f3d[:, rows + G:] = from_hi.numpy()
        f3d[:, :G] = from_lo.numpy()

    f = np.zeros((9, rows + 2 * G, nx), dtype=np.float32)
    f[:, G:G + rows] = f0[:, y0:y0 + rows]
    steps = S * passes
    totals = np.zeros(steps)
    # start of a run: stand-alone acceleration of the owned row, then the ghost refresh
    if y0 <= ny - 2 < y0 + rows:
        _accelerate_row(case, f, ny - 2 - y0 + G, pad_ob[ny - 2 - y0 + G])
    refresh(f)
    t = 0
    for p in range(passes):
        for s_ in range(S):
            nxt, _ = slab.step(np.ascontiguousarray(f.reshape(9, -1)), accel=False)
            totals[t] = slab._speeds.reshape(rows + 2 * G, nx)[G:G + rows].astype(np.float64).sum()
            f = nxt.reshape(9, rows + 2 * G, nx)
            t += 1
            if t < steps:
                accelerate(f)              # folded into the step's store epilogue on the GPU
        refresh(f)

    gathered = [torch.zeros(steps, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(totals))
    np.save(os.path.join(out_dir, "slab%d.npy" % rank), f[:, G:G + rows])
    if rank == 0:
        np.save(os.path.join(out_dir, "av.npy"), sum(g.numpy() for g in gathered) / case.tot_cells)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny,S", [(2, 40, 36, 2), (3, 32, 29, 3), (2, 24, 21, 4)])
def test_stream_pass_protocol_matches_whole_grid_oracle(world, nx, ny, S, tmp_path, lbm):
    import torch.multiprocessing as mp
    from oracle_bindings import Oracle
    from tools import cases
    passes = 7
    mp.spawn(_worker_stream, args=(world, _free_port(), nx, ny, S, passes, str(tmp_path)), nprocs=world, join=True)
    case = cases.random_case(nx, ny, seed=9, walls=False, fill=0.05)
    f = cases.perturbed_state(case, seed=9)
    av = Oracle("f32b200", case).run(f, S * passes)
    got = np.concatenate([np.load(tmp_path / ("slab%d.npy" % r)) for r in range(world)], axis=1)
    assert np.array_equal(got.reshape(9, -1).view(np.uint32), f.view(np.uint32))
    assert np.max(np.abs(np.load(tmp_path / "av.npy") - av) / av) <= 1e-12
