import importlib, os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(1, "/root/repo/tests")
from tools import cases
lbm = importlib.import_module("hpc-lattice-boltzmann_b200")
case = cases.random_case(256, 96, seed=7, walls=True)
f0 = cases.perturbed_state(case, seed=7)
os.environ["LBM_FUSE"] = "2"
for dbg in ("2", "0", "0"):
    os.environ["LBM_STREAM_DEBUG"] = dbg
    try:
        with lbm.Lattice(case.nx, case.ny, case.density, case.accel, case.omega, case.obstacles) as lat:
            lat.upload(f0)
            lat.run(2, f64=True)
            print("debug", dbg, "ok", lat.config, flush=True)
    except Exception as e:
        print("debug", dbg, "FAILED", e, flush=True)
        break
