"""Input cases in the reference's formats: the four shipped ones (rebuilt from the compact
fixtures in tests/golden/) and the seeded synthetic channels named in BASELINE.json.

File formats (reference d2q9-bgk.c): params = 7 scalars one per line, nx ny maxIters
reynolds_dim density accel omega (:499-525); obstacles = "x y 1" per blocked cell (:615-628).
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
SHIPPED = ["128x128", "128x256", "256x256", "1024x1024"]


class Case:
    def __init__(self, name, nx, ny, max_iters, reynolds_dim, density, accel, omega, obstacles):
        self.name = name
        self.nx, self.ny = int(nx), int(ny)
        self.max_iters, self.reynolds_dim = int(max_iters), int(reynolds_dim)
        self.density, self.accel, self.omega = float(density), float(accel), float(omega)
        self.obstacles = np.ascontiguousarray(obstacles, dtype=np.int32).reshape(self.ny, self.nx)

    @property
    def cells(self):
        return self.nx * self.ny

    @property
    def tot_cells(self):
        return int(self.cells - np.count_nonzero(self.obstacles))

    def initial_state(self):
        """float32 planes [9, ny*nx] of d2q9-bgk.c:573-594"""
        d = np.float32(self.density)
        w0 = np.float32(np.float64(d) * 4.0 / 9.0)
        w1 = np.float32(np.float64(d) / 9.0)
        w2 = np.float32(np.float64(d) / 36.0)
        f = np.empty((9, self.cells), dtype=np.float32)
        f[0], f[1:5], f[5:9] = w0, w1, w2
        return f

    def write(self, outdir, iters=None):
        """-> (paramfile, obstaclefile) in the reference's text formats"""
        os.makedirs(outdir, exist_ok=True)
        pf = os.path.join(outdir, "input_%s.params" % self.name)
        of = os.path.join(outdir, "obstacles_%s.dat" % self.name)
        with open(pf, "w") as f:
            f.write("%d\n%d\n%d\n%d\n%r\n%r\n%r\n" % (self.nx, self.ny,
                                                     self.max_iters if iters is None else iters,
                                                     self.reynolds_dim, self.density, self.accel,
                                                     self.omega))
        ys, xs = np.nonzero(self.obstacles)
        with open(of, "w") as f:
            f.write("".join("%d %d 1\n" % (x, y) for x, y in zip(xs.tolist(), ys.tolist())))
        return pf, of


def shipped(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    nx, ny = int(z["nx"]), int(z["ny"])
    ob = np.unpackbits(z["obstacles_packed"])[: nx * ny].reshape(ny, nx)
    return Case(name, nx, ny, int(z["max_iters"]), int(z["reynolds_dim"]), float(z["density"]),
                float(z["accel"]), float(z["omega"]), ob)


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def channel(nx, ny, seed=42, block=8, max_iters=200, rows=None, accel=0.005):
    """Synthetic channel of BASELINE.json configs[3:5]: x-periodic, rows 0 and ny-1 fully blocked,
    plus nx*ny/4096 axis-aligned block x block obstacles at uniform positions in rows
    [8, ny-16) from a fixed-seed 64-bit PRNG (~1.6 % blocked).  `rows=(y0, n)` returns only that
    row range of the same global map (a rank's slab) without materialising the rest."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nblocks = (nx * ny) // 4096 if ny >= 64 and nx >= block else 0
    bx = rng.integers(0, nx - block + 1, size=nblocks)
    by = rng.integers(8, ny - 16 - block + 1, size=nblocks) if nblocks else bx
    y0, n = (0, ny) if rows is None else rows
    ob = np.zeros((n, nx), dtype=np.int32)
    if y0 == 0:
        ob[0, :] = 1
    if y0 + n == ny:
        ob[n - 1, :] = 1
    sel = (by + block > y0) & (by < y0 + n)
    for x, y in zip(bx[sel].tolist(), by[sel].tolist()):
        ya, yb = max(y, y0) - y0, min(y + block, y0 + n) - y0
        ob[ya:yb, x:x + block] = 1
    if rows is not None:
        return ob
    return Case("%dx%d" % (nx, ny), nx, ny, max_iters, 10, 0.1, accel, 1.85, ob)


def random_case(nx, ny, seed, fill=0.08, walls=False, accel=0.01, omega=1.7):
    """small randomized test lattice: scattered obstacles, including on row ny-2 and on the
    periodic seams (x = 0, nx-1; y = 0, ny-1)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    ob = (rng.random((ny, nx)) < fill).astype(np.int32)
    if walls:
        ob[0, :] = ob[-1, :] = 1
    return Case("rand%dx%d_s%d" % (nx, ny, seed), nx, ny, 0, 10, 0.1, accel, omega, ob)


def perturbed_state(case, seed, amp=0.05):
    """equilibrium state with a smooth + random positive perturbation (exercises every branch
    of the accelerate mask and gives non-trivial velocities from step 0)"""
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    f = case.initial_state().astype(np.float64)
    f *= 1.0 + amp * (rng.random(f.shape) - 0.5)
    return f.astype(np.float32)


def _main():
    """python tools/cases.py <shipped-name | NXxNY> <outdir> [--iters N]
    writes input_<name>.params / obstacles_<name>.dat in the reference's text formats: one of the
    four shipped cases (rebuilt from tests/golden/) or the seeded synthetic channel of that size."""
    import argparse
    ap = argparse.ArgumentParser(description=_main.__doc__)
    ap.add_argument("case")
    ap.add_argument("outdir")
    ap.add_argument("--iters", type=int, default=None)
    a = ap.parse_args()
    if a.case in SHIPPED:
        c = shipped(a.case)
    else:
        nx, ny = (int(v) for v in a.case.lower().split("x"))
        c = channel(nx, ny)
    pf, of = c.write(a.outdir, iters=a.iters)
    print(pf)
    print(of)


if __name__ == "__main__":
    _main()
