#!/usr/bin/env python3
"""Small driver for ncu / tuning runs: one lattice, a few timesteps, optional config sweep.

  profile_target.py --workload 16384x4096 --steps 6                     (target for ncu)
  profile_target.py --workload 16384x16384 --steps 100 --sweep          (TPB/VEC/pad sweep)
Environment knobs of the engine (read at lbm_create): LBM_TPB, LBM_VEC, LBM_CHUNK, LBM_GRAPH,
LBM_PLANE_PAD.
"""
import argparse
import importlib
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import cases  # noqa: E402


def run_once(lbm, nx, ny, steps, warmup, shipped=None):
    if shipped:
        c = cases.shipped(shipped)
        ob, dens, acc, om = c.obstacles, c.density, c.accel, c.omega
    else:
        ob, dens, acc, om = cases.channel(nx, ny, rows=(0, ny)), 0.1, 0.005, 1.85
    with lbm.Lattice(nx, ny, dens, acc, om, ob) as lat:
        lat.init_equilibrium()
        lat.run(warmup)
        best = None
        for _ in range(3):
            lat.run(steps)
            ms = lat.last_run_ms
            best = ms if best is None else min(best, ms)
        return {"config": lat.config, "ms_per_step": best / steps,
                "mlups": nx * ny * steps / (best / 1e3) / 1e6,
                "gbs": 72.0 * nx * ny * steps / (best / 1e3) / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="16384x4096")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--shipped", action="store_true", help="workload names a shipped case")
    ap.add_argument("--knobs", default="", help="extra sweep axes, e.g. LBM_FOO=0,1;LBM_BAR=2,3")
    a = ap.parse_args()
    nx, ny = (int(v) for v in a.workload.split("x"))
    lbm = importlib.import_module("hpc-lattice-boltzmann_b200")
    ship = a.workload if a.shipped else None
    if not a.sweep:
        print(json.dumps(run_once(lbm, nx, ny, a.steps, a.warmup, ship)))
        return
    axes = {"LBM_TPB": ["128", "256", "512"], "LBM_VEC": ["4", "2"], "LBM_PLANE_PAD": ["0", "4096"]}
    if a.knobs:
        axes = {}
        for part in a.knobs.split(";"):
            k, v = part.split("=")
            axes[k] = v.split(",")
    keys = list(axes)
    for combo in itertools.product(*[axes[k] for k in keys]):
        for k, v in zip(keys, combo):
            os.environ[k] = v
        r = run_once(lbm, nx, ny, a.steps, a.warmup, ship)
        r["knobs"] = dict(zip(keys, combo))
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
