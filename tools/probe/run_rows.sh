#!/bin/bash
# persistent kernel: rows per block sweep (more rows = fewer SMs, but interior rows hide the halo flight time)
fmt='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d.get("knobs"), "us/step", round(d["ms_per_step"]*1e3,3), "MLUPS", round(d["mlups"]), d["config"][45:125])
    except Exception: print(l[:300])
'
for w in 128x128 128x256 256x256; do
  echo "== $w"
  LBM_RESIDENT=1 timeout 300 python tools/profile_target.py --workload $w --shipped --steps 20000 --warmup 2000 --sweep --knobs "LBM_RES_ROWS=1,2,3,4,6" 2>&1 | python -c "$fmt"
done
for w in 256x512 512x512 512x256; do
  echo "== $w channel"
  LBM_RESIDENT=1 timeout 300 python tools/profile_target.py --workload $w --steps 10000 --warmup 1000 --sweep --knobs "LBM_RES_ROWS=2,3,4,6,8" 2>&1 | python -c "$fmt"
done
