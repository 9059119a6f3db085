"""tuning aid: per-phase clock64 trace of lbm_resident_kernel.  Needs a trace build of the library:
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fopenmp -DLBM_RES_TRACE \
       -shared hpc-lattice-boltzmann_b200/csrc/lbm_engine.cu -o tools/probe/liblbm_trace.so -lcudart
  python tools/probe/trace_res.py 128x128 out.txt"""
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
