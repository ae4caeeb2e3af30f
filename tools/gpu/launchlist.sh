#!/bin/bash
mkdir -p gpurun_out
export CPT_B200_GRAPHS=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.DictReader(l for l in open('gpurun_out/launches.csv') if l.startswith('"')))
# last forward = last 91-ish launches; aggregate the final 'steps' forward: take launches after the last embed_text_ln
idx=[i for i,r in enumerate(rows) if 'embed_text_ln' in r['Kernel Name']]
start=idx[-1]-1
agg=collections.OrderedDict()
for r in rows[start:]:
    n=r['Kernel Name'].split('(')[0].replace('void ','').replace('cptk::','')
    agg.setdefault(n,[0,0.0]); agg[n][0]+=1; agg[n][1]+=float(r['Metric Value'])/1e3
tot=sum(v[1] for v in agg.values())
for n,(c,t) in agg.items(): print('%-70s x%3d %8.1f us total %6.1f us each %5.1f%%'%(n[:70],c,t,t/c,100*t/tot))
print('sum %.1f us'%tot)
PY
