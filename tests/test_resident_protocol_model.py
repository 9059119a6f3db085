"""CPU model of lbm_resident_kernel's block-to-block protocol (hpc-lattice-boltzmann_b200/csrc/lbm_resident.cuh),
run with one Python thread per block and randomised thread switches:

  * block b owns R whole rows (the last block possibly fewer) and keeps them for a whole "launch";
  * per timestep it relaxes its rows (the f32-strict oracle on the block's rows padded with one row
    below and above), then stores the three populations its vertical neighbours pull (4,7,8 of its first
    row downwards, 2,5,6 of its last row upwards) as 64-bit words {step tag : float bits} into the
    neighbours' inboxes -- slot = step parity -- and polls its own inboxes for words carrying the tag of
    the step it has just finished; no other synchronisation exists between blocks;
  * the last step of a launch sends nothing; the next launch starts from the assembled state with tags
    that continue where the previous launch stopped (`base`), so stale words can never match.

What this pins without a GPU: two parity slots suffice (a block is never more than one step ahead of a
neighbour, whatever the interleaving), a block that is its own neighbour on both sides (one block) or whose
two neighbours are the same block (two blocks) works, tags stay unambiguous across launches, and the
assembled lattice equals the whole-grid oracle bit for bit, with the per-step speed sums combined in block
order."""
import os
import random
import sys
import threading

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))

LO_PLANES, HI_PLANES = (4, 7, 8), (2, 5, 6)


def _pack(values, tag):
    return (np.uint64(tag) << np.uint64(32)) | values.view(np.uint32).astype(np.uint64)


def _accelerate_row(case, f3d, row, obst_row):
    """kernels.cl:17-41 on one row of a [9, rows, nx] float32 array"""
    d, a = np.float32(case.density), np.float32(case.accel)
    a1 = np.float32(np.float64(d * a) / 9.0)
    a2 = np.float32(np.float64(d * a) / 36.0)
    r = f3d[:, row, :]
    m = (obst_row == 0) & ((r[3] - a1) > 0) & ((r[6] - a2) > 0) & ((r[7] - a2) > 0)
    r[1][m] += a1; r[5][m] += a2; r[8][m] += a2
    r[3][m] -= a1; r[6][m] -= a2; r[7][m] -= a2


def _launch(case, state, R, nsteps, base, fuse_after, inbox, totals, t_first, seed):
    """one cooperative launch: `nsteps` timesteps of every block, state [9, ny, nx] updated in place"""
    from oracle_bindings import Oracle
    from tools import cases
    nx, ny = case.nx, case.ny
    nblk = -(-ny // R)
    src = state.copy()                      # what every block reads at launch start
    failures = []

    def block(b):
        try:
            rng = random.Random(seed * 1000 + b)
            y0 = b * R
            nr = min(R, ny - y0)
            below, above = (b - 1) % nblk, (b + 1) % nblk
            pad_ob = np.zeros((nr + 2, nx), dtype=np.int32)
            pad_ob[1:-1] = case.obstacles[y0:y0 + nr]
            o = Oracle("f32b200", cases.Case("block", nx, nr + 2, 0, 10, case.density, case.accel, case.omega, pad_ob))
            f = np.zeros((9, nr + 2, nx), dtype=np.float32)
            f[:, 1:-1] = src[:, y0:y0 + nr]
            f[list(HI_PLANES), 0] = src[list(HI_PLANES), (y0 - 1) % ny]          # the row below, from global memory
            f[list(LO_PLANES), nr + 1] = src[list(LO_PLANES), (y0 + nr) % ny]    # the row above
            acc_row = ny - 2 - y0 + 1 if y0 <= ny - 2 < y0 + nr else None
            for t in range(nsteps):
                last = t + 1 == nsteps
                tag, par = base + t + 1, (t + 1) & 1
                nxt, _ = o.step(np.ascontiguousarray(f.reshape(9, -1)), accel=False)
                totals[t_first + t, b] = o._speeds.reshape(nr + 2, nx)[1:-1].astype(np.float64).sum()
                f = nxt.reshape(9, nr + 2, nx)
                if acc_row is not None and (not last or fuse_after):
                    _accelerate_row(case, f, acc_row, pad_ob[acc_row])           # the next step's accelerate_flow
                if last:
                    break
                # halo words out: column by column in a random order, with thread switches in between
                cols = list(range(nx))
                rng.shuffle(cols)
                for x in cols:
                    for j in range(3):
                        inbox[above, 0, par, j, x] = _pack(f[HI_PLANES[j], nr, x:x + 1], tag)[0]
                        inbox[below, 1, par, j, x] = _pack(f[LO_PLANES[j], 1, x:x + 1], tag)[0]
                    if rng.random() < 0.2:
                        threading.Event().wait(rng.random() * 1e-4)
                # halo words in: poll until every word of my two inboxes carries this step's tag
                spins = 0
                while True:
                    mine = inbox[b, :, par]                                      # [2, 3, nx]
                    if np.all((mine >> np.uint64(32)) == np.uint64(tag)):
                        break
                    spins += 1
                    if spins > 200000:
                        raise RuntimeError("block %d timed out at step %d" % (b, t))
                    threading.Event().wait(1e-5)
                words = inbox[b, :, par].copy()
                vals = (words & np.uint64(0xffffffff)).astype(np.uint32).view(np.float32)
                for j in range(3):
                    f[HI_PLANES[j], 0] = vals[0, j]
                    f[LO_PLANES[j], nr + 1] = vals[1, j]
            state[:, y0:y0 + nr] = f[:, 1:-1]
        except Exception as e:  # noqa: BLE001 -- reported by the test thread
            failures.append(e)

    threads = [threading.Thread(target=block, args=(b,)) for b in range(nblk)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if failures:
        raise failures[0]


@pytest.mark.parametrize("nx,ny,R,launches", [(24, 19, 1, (7, 2, 6)), (16, 23, 4, (5, 8)), (20, 9, 9, (6, 3)),
                                              (12, 10, 5, (4, 5))])
def test_halo_word_protocol_matches_whole_grid_oracle(nx, ny, R, launches):
    from oracle_bindings import Oracle
    from tools import cases
    case = cases.random_case(nx, ny, seed=nx * ny + R, walls=False, fill=0.05)     # flow crosses the periodic seam
    f0 = cases.perturbed_state(case, seed=R)
    steps = sum(launches)
    want = f0.copy()
    av = Oracle("f32b200", case).run(want, steps)

    state = f0.reshape(9, ny, nx).copy()
    nblk = -(-ny // R)
    inbox = np.zeros((nblk, 2, 2, 3, nx), dtype=np.uint64)
    totals = np.zeros((steps, nblk))
    # start of a run: the stand-alone accelerate_flow of the first step (the engine's accelerate_row_kernel)
    _accelerate_row(case, state, ny - 2, case.obstacles[ny - 2])
    base, done = 0, 0
    for i, n in enumerate(launches):
        _launch(case, state, R, n, base, i + 1 < len(launches), inbox, totals, done, seed=i)
        base += n
        done += n
    assert np.array_equal(state.reshape(9, -1).view(np.uint32), want.view(np.uint32))
    got_av = totals.sum(axis=1) / case.tot_cells
    assert np.max(np.abs(got_av - av) / av) <= 1e-12
