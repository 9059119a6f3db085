#!/usr/bin/env python3
"""Parity check of the one-process-per-GPU path; launch with torch.distributed.run:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29611 tools/multirank_check.py --nx 256 --ny 96 --steps 40

Every rank owns one row slab on its own GPU (lbm_create_rank), uploads its part of a seeded state,
runs, downloads; rank 0 assembles the lattice and compares it BIT FOR BIT with the f32-strict CPU
oracle run on the whole grid, and the av_vels to 1e-12.  Exit code 0 = parity.
torch.distributed (gloo) is plumbing only: it carries the ncclUniqueId and the gathered slabs.
"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
from tools import cases  # noqa: E402


def main():
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--ny", type=int, default=96)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--channel", action="store_true", help="synthetic channel instead of random obstacles")
    ap.add_argument("--expect-timeout", action="store_true",
                    help="negative test (LBM_TEST_RING_STALL): every rank's run must fail with the time-out error")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    lbm = importlib.import_module("hpc-lattice-boltzmann_b200")

    box = [lbm.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]

    if a.channel:
        case = cases.channel(a.nx, a.ny, accel=0.01)
    else:
        case = cases.random_case(a.nx, a.ny, seed=a.seed, walls=True)
    f0 = cases.perturbed_state(case, seed=a.seed)
    nx = case.nx
    y0, rows = lbm.slab_rows(case.ny, world, rank)
    lat = lbm.Lattice(case.nx, case.ny, case.density, case.accel, case.omega,
                      case.obstacles[y0:y0 + rows], rank=rank, world=world, device=local, unique_id=uid)
    assert (lat.y0, lat.rows) == (y0, rows)
    assert lat.tot_cells == case.tot_cells, (lat.tot_cells, case.tot_cells)
    lat.upload(np.ascontiguousarray(f0[:, y0 * nx:(y0 + rows) * nx]))
    if a.expect_timeout:
        try:
            lat.run(a.steps, f64=True)
            msg = "no error"
        except lbm.LbmError as e:
            msg = str(e)
        lat.close()
        msgs = [None] * world if rank == 0 else None
        dist.gather_object(msg, msgs, dst=0)
        ok = True
        if rank == 0:
            ok = all("timed out waiting for a neighbour GPU" in m for m in msgs)
            print("multirank_check: expected time-out: %s -> %s" % (msgs, "OK" if ok else "FAIL"), flush=True)
        flag = [ok]
        dist.broadcast_object_list(flag, src=0)
        dist.destroy_process_group()
        return 0 if flag[0] else 1
    av1 = lat.run(a.steps, f64=True)
    av2 = lat.run(5, f64=True)                     # a second run continues from the canonical state
    av3 = np.array([lat.step() for _ in range(3)])
    f_local = lat.download()
    avv = float(lat.av_velocity())
    cfg = lat.config
    ms = lat.last_run_ms
    lat.close()

    parts = [None] * world if rank == 0 else None
    dist.gather_object((y0, rows, f_local, av1, av2, av3, avv), parts, dst=0)
    ok = True
    if rank == 0:
        from oracle_bindings import Oracle
        f_gpu = np.empty_like(f0)
        for (py0, prows, pf, *_rest) in parts:
            f_gpu[:, py0 * nx:(py0 + prows) * nx] = pf
        o = Oracle("f32b200", case)
        f = f0.copy()
        r1 = o.run(f, a.steps)
        r2 = o.run(f, 5)
        r3 = o.run(f, 3)
        same = f_gpu.view(np.uint32) == f.view(np.uint32)
        rel = max(np.max(np.abs(av1 - r1) / np.abs(r1)), np.max(np.abs(av2 - r2) / np.abs(r2)))
        rel3 = np.max(np.abs(av3.astype(np.float64) - r3) / np.abs(r3))
        all_same_av = all(np.array_equal(p[3], av1) and np.array_equal(p[4], av2) for p in parts)
        ok = bool(same.all() and rel <= 1e-12 and rel3 <= 1e-6 and all_same_av)
        print("multirank_check: world=%d %dx%d steps=%d [%s] state_bit_exact=%s (%d differ) "
              "av_rel=%.2e ranks_agree=%s -> %s"
              % (world, case.nx, case.ny, a.steps, cfg, bool(same.all()),
                 int(np.count_nonzero(~same)), rel, all_same_av, "OK" if ok else "FAIL"), flush=True)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    return 0 if flag[0] else 1


if __name__ == "__main__":
    sys.exit(main())
