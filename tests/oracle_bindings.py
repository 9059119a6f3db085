"""ctypes bindings of the CPU checkers under oracle/ -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import
this module; nothing on the product path does.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_lbm.so")
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "d2q9-bgk-ref")
CANON_EXE = os.path.join(ROOT, "oracle", "canon")

_VARIANTS = {"f64": (np.float64, C.c_double), "f32ref": (np.float32, C.c_float),
             "f32b200": (np.float32, C.c_float)}
_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.isfile(ORACLE_LIB):
            raise RuntimeError("oracle/liboracle.so not built: run `make -C oracle`")
        _oracle = C.CDLL(ORACLE_LIB)
    return _oracle


class Oracle:
    """one arithmetic variant of oracle/canon_impl.h over a fixed lattice"""

    def __init__(self, variant, case):
        self.np_t, self.c_t = _VARIANTS[variant]
        self.v, self.case, self.lib = variant, case, oracle_lib()
        self.obst = np.ascontiguousarray(case.obstacles, dtype=np.int32).ravel()
        self.n = case.cells
        self.tot_cells = int(self.n - np.count_nonzero(self.obst))
        self._speeds = np.empty(self.n, dtype=self.np_t)

    def _fn(self, name, restype=None):
        fn = getattr(self.lib, "%s_canon_%s" % (self.v, name))
        fn.restype = restype
        return fn

    def _p(self, a):
        return a.ctypes.data_as(C.c_void_p)

    def _real(self, x):
        return self.c_t(float(self.np_t(x)))

    def init(self):
        f = np.empty((9, self.n), dtype=self.np_t)
        self._fn("init")(self.case.nx, self.case.ny, self._real(self.case.density), self._p(f))
        return f

    def step(self, src, accel=True):
        """-> (dst, av_vel as float64); src is modified in place by the acceleration"""
        assert src.dtype == self.np_t and src.flags.c_contiguous
        dst = np.empty_like(src)
        c = self.case
        av = self._fn("step", C.c_double)(c.nx, c.ny, self._real(c.density), self._real(c.accel),
                                          self._real(c.omega), self._p(self.obst), self._p(src),
                                          self._p(dst), self._p(self._speeds),
                                          C.c_long(self.tot_cells), int(accel))
        return dst, av

    def run(self, f, iters):
        """in place; -> av_vels float64[iters]"""
        assert f.dtype == self.np_t and f.flags.c_contiguous
        av = np.empty(max(iters, 1), dtype=np.float64)
        c = self.case
        rc = self._fn("run", C.c_int)(c.nx, c.ny, self._real(c.density), self._real(c.accel),
                                      self._real(c.omega), self._p(self.obst), self._p(f), iters,
                                      self._p(av))
        assert rc == 0
        return av[:iters]

    def av_velocity(self, f):
        return self._fn("av_velocity", C.c_double)(self.case.nx, self.case.ny, self._p(self.obst),
                                                   self._p(f))

    def macroscopic(self, f):
        out = np.empty((4, self.n), dtype=self.np_t)
        self._fn("macroscopic")(self.case.nx, self.case.ny, self._real(self.case.density),
                                self._p(self.obst), self._p(f), self._p(out[0]), self._p(out[1]),
                                self._p(out[2]), self._p(out[3]))
        return out


class Reference:
    """the UNMODIFIED reference (oracle/_ref/libref_lbm.so: d2q9-bgk.c + kernels.cl on the
    host-memory OpenCL shim), driven through its own initialise()/timestep().  Square grids only."""

    def __init__(self, paramfile, obstaclefile, workdir):
        if not os.path.isfile(REF_LIB):
            raise RuntimeError("oracle/_ref/libref_lbm.so not built (needs the reference sources)")
        self.lib = C.CDLL(REF_LIB)
        # the reference's initialise() prints its OpenCL device list to stdout: keep it out of
        # the caller's stdout (bench.py must print exactly one JSON line)
        sys.stdout.flush()
        saved = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            rc = self.lib.ref_open(paramfile.encode(), obstaclefile.encode(), workdir.encode())
            C.CDLL(None).fflush(None)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
        if rc != 0:
            raise RuntimeError("ref_open failed: %d" % rc)
        nx, ny, it, tc = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.lib.ref_shape(C.byref(nx), C.byref(ny), C.byref(it), C.byref(tc))
        self.nx, self.ny, self.max_iters, self.tot_cells = nx.value, ny.value, it.value, tc.value

    def steps(self, n):
        av = np.empty(max(n, 1), dtype=np.float32)
        rc = self.lib.ref_steps(n, av.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return av[:n]

    def upload(self, planes):
        a = np.ascontiguousarray(planes, dtype=np.float32)
        assert self.lib.ref_upload(a.ctypes.data_as(C.c_void_p)) == 0

    def download(self):
        out = np.empty((9, self.nx * self.ny), dtype=np.float32)
        assert self.lib.ref_download(out.ctypes.data_as(C.c_void_p)) == 0
        return out

    def close(self):
        if self.lib is not None:
            self.lib.ref_close()
            self.lib = None
