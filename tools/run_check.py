#!/usr/bin/env python3
"""Run the reference's result checker on a pair of output files.

Two modes, same command line as the reference's check/check.py (check.py:26-57):

  --ref-av-vels-file R1 --ref-final-state-file R2 --av-vels-file A --final-state-file F
  [--tolerance PCT]

1. "reference" mode (default when the reference checkout is present, or --checker PATH):
   executes the reference's check.py UNMODIFIED FROM DISK.  That script is Python-2-only
   (version gate check.py:6-10, print statements), and there is no python2 here, so it is
   loaded as text, its `print X` statements are rewritten to calls in memory, and it is exec'd
   with sys.version_info reporting 2.7.  Nothing is written back; the reference file is not
   copied into this repository.

2. "native" mode (--native, or automatically when no reference checker is reachable, e.g. on
   the GPU box): a Python 3 restatement of the same comparison (check.py:59-147): relative
   difference 100*(ref-sim)/sim on every av_vels entry and every final_state pressure, worst
   entry by absolute percentage, failure if it is non-finite or above the tolerance (1 %).
   Reference files may be the text goldens or the compact .npz fixtures of tests/golden/.
   tests/test_check_tool.py keeps the two modes in agreement.
"""
import argparse
import os
import re
import sys

import numpy as np

DEFAULT_CHECKER = "/root/reference/check/check.py"


def run_reference_checker(checker_path, argv):
    """exec the on-disk py2 checker under py3; returns its exit code"""
    src = open(checker_path).read()
    out = []
    for line in src.split("\n"):
        m = re.match(r"^(\s*)print\s+(.+)$", line)
        if m:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        elif re.match(r"^\s*print\s*$", line):
            line = line.replace("print", "print()")
        out.append(line)
    code = compile("\n".join(out), checker_path, "exec")

    class _V(tuple):
        major, minor = 2, 7

    real_vi, real_argv = sys.version_info, sys.argv
    sys.version_info = _V((2, 7, 18, "final", 0))
    sys.argv = [checker_path] + list(argv)
    try:
        exec(code, {"__name__": "__main__", "__file__": checker_path, "exit": sys.exit})
        rc = 0
    except SystemExit as e:
        rc = e.code if isinstance(e.code, int) else (0 if e.code is None else 1)
    finally:
        sys.version_info, sys.argv = real_vi, real_argv
    return rc


def load_av_vels(path):
    if path.endswith(".npz"):
        return np.asarray(np.load(path)["av_vels"], dtype=np.float64)
    return np.loadtxt(path, usecols=[1])


def load_final_state(path):
    """-> (coords [n,2], pressure [n], selection or None).  A subsampled fixture carries the
    flat indices it kept; the simulated file is reduced to the same cells."""
    if path.endswith(".npz"):
        z = np.load(path)
        nx, ny = int(z["nx"]), int(z["ny"])
        idx = z["pressure_index"] if "pressure_index" in z.files else np.arange(nx * ny)
        coords = np.stack([idx % nx, idx // nx], axis=1).astype(np.float64)
        return coords, np.asarray(z["pressure"], dtype=np.float64), idx
    a = np.loadtxt(path, usecols=[0, 1, 5])
    return a[:, 0:2], a[:, 2], None


def diff_values(ref, sim):
    with np.errstate(divide="ignore", invalid="ignore"):
        diff = ref - sim
        pct = 100.0 * (diff / (ref - diff))
    worst = int(np.argmax(np.abs(pct)))
    return dict(max_diff_step=worst, max_diff=diff[worst], max_diff_pcnt=pct[worst],
                sim_val=sim[worst], ref_val=ref[worst], total=float(np.sum(np.abs(diff))))


def native_check(ref_av, ref_fs, sim_av, sim_fs, tolerance, quiet=False):
    say = (lambda *a: None) if quiet else print
    av_ref = load_av_vels(ref_av)
    av_sim = load_av_vels(sim_av)
    c_ref, p_ref, sel = load_final_state(ref_fs)
    c_sim, p_sim, _ = load_final_state(sim_fs)
    if sel is not None and len(p_sim) != len(p_ref):
        c_sim, p_sim = c_sim[sel], p_sim[sel]
    if c_ref.shape != c_sim.shape or np.any(c_ref != c_sim):
        say("Final state files coordinates were not the same")
        return 1, None
    if av_ref.size != av_sim.size:
        say("Different number of steps in av_vels files")
        return 1, None
    a = diff_values(av_ref, av_sim)
    say("Total difference in av_vels : {total:.12E}".format(**a))
    say("Biggest difference (at step {max_diff_step:d}) : {max_diff:.12E}".format(**a))
    say("  {sim_val:.12E} vs. {ref_val:.12E} = {max_diff_pcnt:.2g}%".format(**a))
    say()
    f = diff_values(p_ref, p_sim)
    f["jj"], f["ii"] = int(c_sim[f["max_diff_step"], 0]), int(c_sim[f["max_diff_step"], 1])
    say("Total difference in final_state : {total:.12E}".format(**f))
    say("Biggest difference (at coord ({jj:d},{ii:d})) : {max_diff:.12E}".format(**f))
    say("  {sim_val:.12E} vs. {ref_val:.12E} = {max_diff_pcnt:.2g}%".format(**f))
    say()
    fs_failed = (not np.isfinite(f["max_diff_pcnt"])) or abs(f["max_diff_pcnt"]) > tolerance
    av_failed = (not np.isfinite(a["max_diff_pcnt"])) or abs(a["max_diff_pcnt"]) > tolerance
    if fs_failed:
        say("final state failed check")
    if av_failed:
        say("av_vels failed check")
    if not (fs_failed or av_failed):
        say("Both tests passed!")
    return (1 if (fs_failed or av_failed) else 0), dict(av_vels=a, final_state=f)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    ap.add_argument("--tolerance", type=float, default=1.0)
    ap.add_argument("--ref-av-vels-file", required=True)
    ap.add_argument("--ref-final-state-file", required=True)
    ap.add_argument("--av-vels-file", required=True)
    ap.add_argument("--final-state-file", required=True)
    ap.add_argument("--native", action="store_true", help="use the Python 3 restatement")
    ap.add_argument("--checker", default=DEFAULT_CHECKER, help="path of the reference check.py")
    a = ap.parse_args(argv)
    npz = a.ref_av_vels_file.endswith(".npz") or a.ref_final_state_file.endswith(".npz")
    if not a.native and not npz and os.path.isfile(a.checker):
        passthru = ["--tolerance", repr(a.tolerance),
                    "--ref-av-vels-file", a.ref_av_vels_file,
                    "--ref-final-state-file", a.ref_final_state_file,
                    "--av-vels-file", a.av_vels_file,
                    "--final-state-file", a.final_state_file]
        print("[run_check] executing the reference checker %s unmodified" % a.checker)
        return run_reference_checker(a.checker, passthru)
    print("[run_check] native Python 3 restatement of check.py")
    rc, _ = native_check(a.ref_av_vels_file, a.ref_final_state_file, a.av_vels_file,
                         a.final_state_file, a.tolerance)
    return rc


if __name__ == "__main__":
    sys.exit(main())
