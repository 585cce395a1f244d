#!/bin/bash
# round 2, GPU call 28 (1 GPU): why is the 800x800 frame at 29 ms? A/B of the round renderer (leap on/off, compaction off) + launch lists
mkdir -p gpurun_out
for tag in default noleap nocompact; do
  case $tag in default) E="";; noleap) E="NGP_FRAME_LEAP=0";; nocompact) E="NGP_FRAME_COMPACT=0";; esac
  env $E timeout 300 python bench.py --config frame800 --steps 10 --warmup 3 --cpu-budget 1 > gpurun_out/r2_c28_frame_$tag.json 2>/dev/null; echo "$tag rc=$?"
  grep '^{' gpurun_out/r2_c28_frame_$tag.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$tag', d['value'], d['ms_per_step'], d['config'].get('samples_per_ray'), d['e2e']['ms_per_step'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --profile-from-start off \
    --csv --log-file gpurun_out/r2_c28_launches_frame800.csv python bench.py --config frame800 --ncu-window 1 > gpurun_out/r2_c28_ncu_frame.log 2>&1; echo "ncu frame rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_c28_launches_frame800.csv')))
h=next(i for i,r in enumerate(rows) if 'Kernel Name' in r); idx={n:i for i,n in enumerate(rows[h])}
tot={}
for r in rows[h+1:]:
    if len(r)<=idx['Metric Value']: continue
    try: v=float(r[idx['Metric Value']].replace(',',''))
    except: continue
    k=r[idx['Kernel Name']][:50]; u=r[idx['Metric Unit']]
    v = v/1e3 if u in ('ns','nsecond') else (v*1e3 if u in ('ms','msecond') else v)
    t=tot.setdefault(k,[0,0.0]); t[0]+=1; t[1]+=v
for k,(c,v) in sorted(tot.items(), key=lambda x:-x[1][1])[:12]: print(f'{v:10.1f} us  x{c:3d}  {k}')
PY
