"""World-size-2/3 CPU model (gloo) of the multi-GPU protocol: the same slab arithmetic
(lbm_slab_rows from the C ABI), the same ring-periodic halo (rows' populations 4,7,8 travel down,
2,5,6 travel up), the accelerated row living on whichever rank owns global row ny-2, and the
per-step speed totals combined in rank order.  Each rank advances its padded slab with the
f32-strict oracle; the assembled lattice must equal the oracle run on the whole grid bit for bit.
This pins the host-side decomposition logic that bench.py / tools/multirank_check.py and the
engine's lbm_create_rank share, without a GPU.  A second model does the same for the
two-timesteps-per-pass protocol (ghost depth 1, t+1 boundary rows through 3-row strips, fix-up of rows
1 and `rows`), i.e. what lbm_fused2_kernel and its strip launches do across ranks."""
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


# ---- the two-timesteps-per-pass protocol (what lbm_fused2_kernel + the strip fix-ups do) ---------
def _worker_two_step(rank, world, port, nx, ny, pairs, out_dir):
    """CPU model of one rank: ghost depth stays 1.  Per pass: t+1 on all own rows, its boundary
    populations to the neighbours' strips; t+2 on rows 2..rows-1 from the own t+1 state; rows 1 and
    `rows` of t+2 from the 3-row strips [neighbour row | own boundary row | own next row]; t+2 ghost
    rows exchanged as usual."""
    sys.path.insert(0, ROOT)
    sys.path.insert(1, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle_bindings import Oracle
    from tools import cases
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lbm = importlib.import_module("hpc-lattice-boltzmann_b200")

    case = cases.random_case(nx, ny, seed=9, walls=True)
    f0 = cases.perturbed_state(case, seed=9).reshape(9, ny, nx)
    y0, rows = lbm.slab_rows(ny, world, rank)
    lo, hi = (rank - 1) % world, (rank + 1) % world
    pad_ob = np.zeros((rows + 2, nx), dtype=np.int32)
    pad_ob[1:-1] = case.obstacles[y0:y0 + rows]
    slab = Oracle("f32b200", cases.Case("slab", nx, rows + 2, 0, 10, case.density, case.accel, case.omega, pad_ob))
    owns_accel = y0 <= ny - 2 < y0 + rows
    arow = ny - 2 - y0 + 1

    def strip_oracle(obst_row):
        ob = np.zeros((3, nx), dtype=np.int32)
        ob[1] = obst_row
        return Oracle("f32b200", cases.Case("strip", nx, 3, 0, 10, case.density, case.accel, case.omega, ob))

    fix_lo, fix_hi = strip_oracle(pad_ob[1]), strip_oracle(pad_ob[rows])

    def sendrecv(down, up):
        """down -> lower neighbour, up -> upper neighbour; returns (from_upper, from_lower)"""
        d, u = torch.from_numpy(np.ascontiguousarray(down)), torch.from_numpy(np.ascontiguousarray(up))
        from_hi, from_lo = torch.empty_like(d), torch.empty_like(u)
        reqs = [dist.isend(d, lo, tag=1), dist.isend(u, hi, tag=2),
                dist.irecv(from_hi, hi, tag=1), dist.irecv(from_lo, lo, tag=2)]
        for r in reqs:
            r.wait()
        return from_hi.numpy(), from_lo.numpy()

    def step(o, f3d):
        out, _ = o.step(np.ascontiguousarray(f3d.reshape(9, -1)), accel=False)
        return out.reshape(f3d.shape), o._speeds.reshape(f3d.shape[1:]).astype(np.float64)

    f = np.zeros((9, rows + 2, nx), dtype=np.float32)
    f[:, 1:-1] = f0[:, y0:y0 + rows]
    got_hi, got_lo = sendrecv(f[list(LO_PLANES), 1], f[list(HI_PLANES), rows])       # initial ghosts
    f[list(LO_PLANES), rows + 1], f[list(HI_PLANES), 0] = got_hi, got_lo
    totals = np.zeros(2 * pairs)
    for p in range(pairs):
        if owns_accel:
            _accelerate_row(case, f, arow, pad_ob[arow])
        f1, sp1 = step(slab, f)                                   # t+1 on rows 1..rows
        totals[2 * p] = sp1[1:-1].sum()
        if owns_accel:                                            # step t+2 follows inside the pass
            _accelerate_row(case, f1, arow, pad_ob[arow])
        got_hi, got_lo = sendrecv(f1[list(LO_PLANES), 1], f1[list(HI_PLANES), rows])   # strips' ghost rows
        f2, sp2 = step(slab, f1)                                  # valid on rows 2..rows-1 only
        t2 = sp2[2:rows].sum()
        s_lo = np.zeros((9, 3, nx), dtype=np.float32)             # [lower neighbour's last | my 1 | my 2]
        s_lo[list(HI_PLANES), 0] = got_lo
        s_lo[:, 1], s_lo[:, 2] = f1[:, 1], f1[:, 2]
        s_hi = np.zeros((9, 3, nx), dtype=np.float32)             # [my rows-1 | my rows | upper neighbour's first]
        s_hi[:, 0], s_hi[:, 1] = f1[:, rows - 1], f1[:, rows]
        s_hi[list(LO_PLANES), 2] = got_hi
        r_lo, sp_lo = step(fix_lo, s_lo)
        r_hi, sp_hi = step(fix_hi, s_hi)
        f2[:, 1], f2[:, rows] = r_lo[:, 1], r_hi[:, 1]
        totals[2 * p + 1] = t2 + sp_lo[1].sum() + sp_hi[1].sum()
        got_hi, got_lo = sendrecv(f2[list(LO_PLANES), 1], f2[list(HI_PLANES), rows])   # t+2 ghost rows
        f2[list(LO_PLANES), rows + 1], f2[list(HI_PLANES), 0] = got_hi, got_lo
        f = f2

    gathered = [torch.zeros(2 * pairs, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(totals))
    np.save(os.path.join(out_dir, "slab%d.npy" % rank), f[:, 1:-1])
    if rank == 0:
        np.save(os.path.join(out_dir, "av.npy"), sum(g.numpy() for g in gathered) / case.tot_cells)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny", [(2, 40, 36), (3, 32, 29)])
def test_two_step_pass_protocol_matches_whole_grid_oracle(world, nx, ny, tmp_path, lbm):
    import torch.multiprocessing as mp
    from oracle_bindings import Oracle
    from tools import cases
    pairs = 9
    mp.spawn(_worker_two_step, args=(world, _free_port(), nx, ny, pairs, str(tmp_path)), nprocs=world, join=True)
    case = cases.random_case(nx, ny, seed=9, walls=True)
    f = cases.perturbed_state(case, seed=9)
    av = Oracle("f32b200", case).run(f, 2 * pairs)
    got = np.concatenate([np.load(tmp_path / ("slab%d.npy" % r)) for r in range(world)], axis=1)
    assert np.array_equal(got.reshape(9, -1).view(np.uint32), f.view(np.uint32))
    assert np.max(np.abs(np.load(tmp_path / "av.npy") - av) / av) <= 1e-12
