for H in 126 50; do LBM_TILE_H=$H LBM_STREAM_TRACE=gpurun_out/r2_trace_H$H.csv python tools/profile_target.py --workload 16384x2048 --steps 20 --warmup 6 > /dev/null 2>&1; done
python - <<'PY'
import csv
for H in (126, 50):
    rows=list(csv.DictReader(open('gpurun_out/r2_trace_H%d.csv'%H)))
    st=[int(r['start_ns']) for r in rows]; en=[int(r['end_ns']) for r in rows]
    t0=min(st); T=max(en)-t0
    dur=sorted(e-s for s,e in zip(st,en))
    print("H=%d tiles=%d kernel span %.1f us; tile duration min/med/max %.1f/%.1f/%.1f us" % (H, len(rows), T/1e3, dur[0]/1e3, dur[len(dur)//2]/1e3, dur[-1]/1e3))
    # active CTAs over time (20 bins)
    nb=20
    act=[0.0]*nb
    for s,e in zip(st,en):
        for b in range(nb):
            lo=t0+T*b/nb; hi=t0+T*(b+1)/nb
            ov=max(0,min(e,hi)-max(s,lo))
            act[b]+=ov/(hi-lo)
    print("  avg resident CTAs per 5%% of the span:", [round(a) for a in act])
    # start times of first 300 blocks relative
    byblock=sorted((int(r['block']),int(r['start_ns'])-t0,int(r['end_ns'])-t0) for r in rows)
    print("  first blocks start (us):", [round(b[1]/1e3,1) for b in byblock[:5]], "block 295:", round(byblock[295][1]/1e3,1), "block 296:", round(byblock[296][1]/1e3,1), "block 600:", round(byblock[600][1]/1e3,1))
    print("  last 5 tiles end (us):", [round(x/1e3,1) for x in sorted(e-t0 for e in en)[-5:]])
PY
