"""ctypes view of liblbm_b200.so (the C ABI in include/lbm_b200.h).

The product is the CUDA library + the C host program in this directory; this module exists so
that pytest and bench.py can drive the same entry points the C program uses.  It adds nothing:
every method is one C call.  There is no fallback of any kind -- if the shared library has not
been built (`make -C hpc-lattice-boltzmann_b200`, or __graft_entry__.build()) importing
`load()` raises, and without a GPU `Lattice(...)` raises with the library's error text.

Import with importlib (the directory name carries a hyphen):
    lbm = importlib.import_module("hpc-lattice-boltzmann_b200")
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblbm_b200.so")
EXE_PATH = os.path.join(HERE, "d2q9-bgk.exe")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "lbm_b200.h")


class LbmError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("density", C.c_float), ("accel", C.c_float),
                ("omega", C.c_float)]


_lib = None
_PP = C.POINTER(C.c_float) * 9


def load():
    """dlopen the CUDA library and declare the prototypes; loud failure if it is not built"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise LbmError("%s not built: run `make -C %s` (needs nvcc); there is no fallback path"
                       % (LIB_PATH, HERE))
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    proto = {
        "lbm_create": (C.c_int, [C.POINTER(vp), C.POINTER(Params), ip, C.c_int]),
        "lbm_create_rank": (C.c_int, [C.POINTER(vp), C.POINTER(Params), ip, C.c_int, C.c_int,
                                      C.c_int, vp]),
        "lbm_comm_unique_id": (C.c_int, [vp]),
        "lbm_slab_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, ip, ip]),
        "lbm_destroy": (None, [vp]),
        "lbm_init_equilibrium": (C.c_int, [vp]),
        "lbm_upload": (C.c_int, [vp, _PP]),
        "lbm_download": (C.c_int, [vp, _PP]),
        "lbm_step": (C.c_int, [vp, fp]),
        "lbm_run": (C.c_int, [vp, C.c_int, fp]),
        "lbm_run_f64": (C.c_int, [vp, C.c_int, dp]),
        "lbm_av_velocity": (C.c_int, [vp, fp]),
        "lbm_macroscopic": (C.c_int, [vp, fp, fp, fp, fp]),
        "lbm_total_density": (C.c_int, [vp, dp]),
        "lbm_last_run_ms": (C.c_double, [vp]),
        "lbm_last_run_launches": (C.c_longlong, [vp]),
        "lbm_tot_cells": (C.c_longlong, [vp]),
        "lbm_local_slab": (C.c_int, [vp, ip, ip]),
        "lbm_config_string": (C.c_char_p, [vp]),
        "lbm_probe_l2_copy": (C.c_int, [C.c_ulonglong, C.c_int, dp]),
        "lbm_host_alloc": (C.c_int, [C.POINTER(vp), C.c_ulonglong]),
        "lbm_host_free": (None, [vp]),
        "lbm_last_error": (C.c_char_p, []),
        "lbm_device_count": (C.c_int, []),
    }
    for name, (res, args) in proto.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def exported_symbols():
    """names declared in include/lbm_b200.h (parsed from the header text)"""
    import re
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", text)))


def _check(rc, what):
    if rc != 0:
        raise LbmError("%s failed: %s" % (what, load().lbm_last_error().decode()))


def _planes(a, n):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.shape == (9, n), a.shape
    return a, _PP(*[a[k].ctypes.data_as(C.POINTER(C.c_float)) for k in range(9)])


class PinnedPlanes:
    """9 float32 planes in page-locked host memory (lbm_host_alloc)"""

    def __init__(self, cells):
        self.lib = load()
        self.ptr = C.c_void_p()
        self.nbytes = 9 * cells * 4
        _check(self.lib.lbm_host_alloc(C.byref(self.ptr), self.nbytes), "lbm_host_alloc")
        buf = (C.c_float * (9 * cells)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float32).reshape(9, cells)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.lbm_host_free(self.ptr)
            self.ptr = None


class Lattice:
    """one lbm_lattice handle.  obstacles: int array [ny, nx] (or the slab's rows in rank mode)"""

    def __init__(self, nx, ny, density, accel, omega, obstacles, ngpus=1, rank=None, world=1,
                 device=0, unique_id=None):
        self.lib = load()
        self.nx, self.ny = int(nx), int(ny)
        self.params = Params(self.nx, self.ny, density, accel, omega)
        ob = np.ascontiguousarray(obstacles, dtype=np.int32)
        self.h = C.c_void_p()
        if rank is None:
            assert ob.size == self.nx * self.ny
            _check(self.lib.lbm_create(C.byref(self.h), C.byref(self.params),
                                       ob.ctypes.data_as(C.POINTER(C.c_int)), ngpus), "lbm_create")
        else:
            uid = C.c_char_p(unique_id) if unique_id is not None else None
            _check(self.lib.lbm_create_rank(C.byref(self.h), C.byref(self.params),
                                            ob.ctypes.data_as(C.POINTER(C.c_int)), rank, world,
                                            device, C.cast(uid, C.c_void_p)), "lbm_create_rank")
        y0, rows = C.c_int(), C.c_int()
        _check(self.lib.lbm_local_slab(self.h, C.byref(y0), C.byref(rows)), "lbm_local_slab")
        self.y0, self.rows = y0.value, rows.value
        self.cells = self.rows * self.nx          # cells held in this process's host planes

    def close(self):
        if self.h:
            self.lib.lbm_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def init_equilibrium(self):
        _check(self.lib.lbm_init_equilibrium(self.h), "lbm_init_equilibrium")

    def upload(self, planes):
        a, pp = _planes(planes, self.cells)
        _check(self.lib.lbm_upload(self.h, pp), "lbm_upload")

    def download(self, out=None):
        if out is None:
            out = np.empty((9, self.cells), dtype=np.float32)
        a, pp = _planes(out, self.cells)
        assert a is out or np.shares_memory(a, out)
        _check(self.lib.lbm_download(self.h, pp), "lbm_download")
        return out

    def step(self):
        v = C.c_float()
        _check(self.lib.lbm_step(self.h, C.byref(v)), "lbm_step")
        return np.float32(v.value)

    def run(self, iters, f64=False):
        if f64:
            av = np.empty(max(iters, 1), dtype=np.float64)
            _check(self.lib.lbm_run_f64(self.h, iters, av.ctypes.data_as(C.POINTER(C.c_double))),
                   "lbm_run_f64")
        else:
            av = np.empty(max(iters, 1), dtype=np.float32)
            _check(self.lib.lbm_run(self.h, iters, av.ctypes.data_as(C.POINTER(C.c_float))),
                   "lbm_run")
        return av[:iters]

    def av_velocity(self):
        v = C.c_float()
        _check(self.lib.lbm_av_velocity(self.h, C.byref(v)), "lbm_av_velocity")
        return np.float32(v.value)

    def total_density(self):
        v = C.c_double()
        _check(self.lib.lbm_total_density(self.h, C.byref(v)), "lbm_total_density")
        return v.value

    def macroscopic(self):
        out = np.empty((4, self.cells), dtype=np.float32)
        p = [out[k].ctypes.data_as(C.POINTER(C.c_float)) for k in range(4)]
        _check(self.lib.lbm_macroscopic(self.h, *p), "lbm_macroscopic")
        return out

    @property
    def last_run_ms(self):
        return float(self.lib.lbm_last_run_ms(self.h))

    @property
    def last_run_launches(self):
        return int(self.lib.lbm_last_run_launches(self.h))

    @property
    def tot_cells(self):
        return int(self.lib.lbm_tot_cells(self.h))

    @property
    def config(self):
        return self.lib.lbm_config_string(self.h).decode()


def slab_rows(ny, world, rank):
    y0, rows = C.c_int(), C.c_int()
    _check(load().lbm_slab_rows(ny, world, rank, C.byref(y0), C.byref(rows)), "lbm_slab_rows")
    return y0.value, rows.value


def probe_l2_copy(nbytes=24 << 20, reps=200):
    """GB/s (read + write) of an L2-resident device-to-device copy, one launch"""
    v = C.c_double()
    _check(load().lbm_probe_l2_copy(nbytes, reps, C.byref(v)), "lbm_probe_l2_copy")
    return v.value


def device_count():
    return int(load().lbm_device_count())


def comm_unique_id():
    """128 bytes of an ncclUniqueId (rank 0 makes it, the caller distributes it)"""
    buf = C.create_string_buffer(128)
    _check(load().lbm_comm_unique_id(C.cast(buf, C.c_void_p)), "lbm_comm_unique_id")
    return buf.raw
