timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "two_step or stream or falls_back or seam or nan_hazard or acceleration or full_size" > gpurun_out/r2aa_t1.log 2>&1; tail -2 gpurun_out/r2aa_t1.log
for R in 2048 16384; do for c in 0 7; do LBM_FUSE=2 LBM_STREAM_CFG=$c python tools/profile_target.py --workload 16384x$R --steps 200 --warmup 20 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('R=$R cfg=$c', round(d['mlups']), 'us/pass', round(d['ms_per_step']*2e3,1), d['config'][100:170])
    except Exception: print(l[:300])
"; done; done
for H in 126 50; do LBM_TILE_H=$H python tools/profile_target.py --workload 16384x2048 --steps 200 --warmup 20 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('H=$H', round(d['mlups']), 'us/pass', round(d['ms_per_step']*2e3,1))
    except Exception: print(l[:300])
"; done
