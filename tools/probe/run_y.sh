timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "two_step or stream or falls_back or seam or nan_hazard or acceleration or full_size" > gpurun_out/r2y_t1.log 2>&1; tail -2 gpurun_out/r2y_t1.log
for pdl in 1 0; do LBM_STREAM_PDL=$pdl python tools/profile_target.py --workload 16384x2048 --steps 200 --warmup 20 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(round(d['mlups']), round(d['ms_per_step']*1e3,1), d['config'][60:170])
"; done
python tools/profile_target.py --workload 16384x16384 --steps 100 --warmup 10 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(round(d['mlups']), round(d['ms_per_step']*1e3,1), d['config'][60:170])
"
for c in 0 7 8 9; do LBM_FUSE=2 LBM_STREAM_CFG=$c python tools/profile_target.py --workload 16384x16384 --steps 100 --warmup 10 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print($c, round(d['mlups']), round(d['ms_per_step']*1e3,1), d['config'][60:170])
    except Exception: print(l[:300])
"; done
