"""bench.py's reference arm runs on the CPU, so its JSON contract can be checked without a GPU:
exactly one line on stdout, the keys the driver reads, and the tier's extra objects."""
import json
import os
import subprocess
import sys

import pytest

from tools import cases

REF_LIB = os.path.join(cases.ROOT, "oracle", "_ref", "libref_lbm.so")


@pytest.mark.skipif(not os.path.isfile(REF_LIB), reason="oracle/_ref not built")
def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")      # what torch.distributed.run exports
    r = subprocess.run([sys.executable, os.path.join(cases.ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "3", "--warmup", "1"], capture_output=True, text=True, env=env,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MLUPS" and d["unit"] == "MLUPS"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    # all host threads are used even though the launcher exported OMP_NUM_THREADS=1
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(cases.ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "3"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
