#!/bin/bash
# closing run on one B200: full GPU test-suite, the persistent-kernel tests three more times, smoke, default bench line
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_tests.log 2>&1
tail -3 gpurun_out/r2i_tests.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent or run_bit_exact or shipped_case" 2>&1 | tail -1; done | tee gpurun_out/r2i_tests_repeat.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1
tail -4 gpurun_out/r2i_smoke.log | cut -c1-160
timeout 700 python bench.py > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
tail -c 400 gpurun_out/r2i_bench_n1.err
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2i_bench_n1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "frac_dram", d["roofline"]["frac_dram"], "1024", d["roofline_1024"]["value"], "launches", d["gpu_launches"])
    for k, v in d["extra"]["small_cases"].items():
        print(k, v if "error" in k else (round(v["persistent"]["us_per_step"], 3), round(v["launches"]["us_per_step"], 3), round(v["speedup"], 2)))
except Exception as e:
    print("bench parse failed", e)
P
