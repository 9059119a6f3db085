"""tuning aid: per-phase clock64 trace of lbm_resident_kernel (needs tools/probe/liblbm_trace.so, built with -DLBM_RES_TRACE)"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools import cases
lbm = importlib.import_module("hpc-lattice-boltzmann_b200")
lbm.LIB_PATH = os.path.join(ROOT, "tools", "probe", "liblbm_trace.so")
name = sys.argv[1]
c = cases.shipped(name)
with lbm.Lattice(c.nx, c.ny, c.density, c.accel, c.omega, c.obstacles) as lat:
    print(lat.config)
    lat.init_equilibrium()
    lat.run(1000)
    lat.run(1000)
    os.environ["LBM_RES_TRACE_FILE"] = sys.argv[2]
    lat.run(2)
    print("ms/step", lat.last_run_ms)
