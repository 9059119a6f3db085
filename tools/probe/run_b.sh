for c in 8 9; do LBM_FUSE=2 LBM_STREAM_CFG=$c python tools/profile_target.py --workload 16384x16384 --steps 120 --warmup 12 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('cfg=$c', round(d['mlups']), 'ms/step', round(d['ms_per_step'],4), d['config'][45:170])
    except Exception: print(l[:300])
"; done
LBM_FUSE=2 LBM_STREAM_CFG=8 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "two_step_passes_bit_exact" 2>&1 | tail -2
