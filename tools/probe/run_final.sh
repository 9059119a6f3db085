#!/bin/bash
# round-2 closing run on one B200: full GPU test-suite, smoke, default bench line, ncu of the persistent kernel
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_tests.log 2>&1
tail -4 gpurun_out/r2f_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
tail -4 gpurun_out/r2f_smoke.log | cut -c1-200
timeout 700 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
tail -c 600 gpurun_out/r2f_bench_n1.err
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2f_bench_n1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "frac_dram", d["roofline"]["frac_dram"], "1024", d["roofline_1024"]["value"])
    for k, v in d["extra"]["small_cases"].items():
        print(k, v if "error" in k else (round(v["persistent"]["us_per_step"], 3), round(v["launches"]["us_per_step"], 3), round(v["speedup"], 2)))
except Exception as e:
    print("bench parse failed", e)
P
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lbm_resident -s 1 -c 1 -f -o gpurun_out/r2f_ncu_resident \
  python tools/profile_target.py --workload 256x256 --shipped --steps 200 --warmup 200 > gpurun_out/r2f_ncu_log.txt 2>&1
tail -3 gpurun_out/r2f_ncu_log.txt | cut -c1-300
