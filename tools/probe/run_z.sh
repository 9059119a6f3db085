for H in 126 50; do for R in 1024 2048 4096 8192; do LBM_TILE_H=$H python tools/profile_target.py --workload 16384x$R --steps 200 --warmup 20 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('H=$H R=$R', round(d['mlups']), 'us/pass', round(d['ms_per_step']*2e3,1), d['config'][118:150])
    except Exception: print(l[:300])
"; done; done
