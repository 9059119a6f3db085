for H in 0 14 22 30 46; do for c in 0 5; do LBM_FUSE=2 LBM_STREAM_CFG=$c LBM_TILE_H=$H python tools/profile_target.py --workload 1024x1024 --shipped --steps 4000 --warmup 400 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('H=$H cfg=$c', round(d['mlups']), 'us/step', round(d['ms_per_step']*1e3,2), d['config'][45:150])
    except Exception: print(l[:300])
"; done; done
LBM_FUSE=1 python tools/profile_target.py --workload 1024x1024 --shipped --steps 4000 --warmup 400 2>&1 | cut -c1-200
