#!/bin/bash
# persistent small-lattice kernel: parity subset, then thread-count sweep on small lattices
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent or run_bit_exact or launch_variants" > gpurun_out/r2e_t1.log 2>&1
tail -15 gpurun_out/r2e_t1.log
fmt='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d.get("knobs"), "us/step", round(d["ms_per_step"]*1e3,3), "MLUPS", round(d["mlups"]), d["config"][38:130])
    except Exception: print(l[:300])
'
for w in 128x128 128x256 256x256; do
  echo "== $w launches / persistent"
  LBM_RESIDENT=0 timeout 120 python tools/profile_target.py --workload $w --shipped --steps 20000 --warmup 2000 2>&1 | python -c "$fmt"
  LBM_RESIDENT=1 timeout 300 python tools/profile_target.py --workload $w --shipped --steps 20000 --warmup 2000 --sweep --knobs "LBM_RES_TPB=128,256,512" 2>&1 | python -c "$fmt"
done
for w in 256x512 512x512 64x64 1024x128; do
  echo "== $w channel launches / persistent"
  LBM_RESIDENT=0 timeout 120 python tools/profile_target.py --workload $w --steps 10000 --warmup 1000 2>&1 | python -c "$fmt"
  LBM_RESIDENT=1 timeout 300 python tools/profile_target.py --workload $w --steps 10000 --warmup 1000 --sweep --knobs "LBM_RES_TPB=256,512" 2>&1 | python -c "$fmt"
done
