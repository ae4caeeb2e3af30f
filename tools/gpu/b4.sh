CPT_B200_TRAIN_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/b4_launches.csv python tools/train_bench.py --batch 4 --dropout 0.1 --steps 1 --warmup 2 > gpurun_out/b4.log 2>&1
python - <<'PY'
import csv
from collections import OrderedDict
lines=[l for l in open('gpurun_out/b4_launches.csv') if l.startswith('"')]
rows=list(csv.DictReader(lines))
# take the last third (one steady-state step): find index of last 'refresh_weights_kernel'
names=[r['Kernel Name'] for r in rows]
idx=[i for i,n in enumerate(names) if 'refresh_weights' in n]
start=idx[-2] if len(idx)>=2 else 0
end=idx[-1]
agg=OrderedDict()
for r in rows[start:end]:
    k=r['Kernel Name'].split('(')[0].replace('void ','')[:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r['Metric Value'])/1e3
tot=sum(v[1] for v in agg.values()); n=sum(v[0] for v in agg.values())
for k,(c,us) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]: print('%-62s %4d %8.1f us %6.1f each'%(k,c,us,us/c))
print('launches',n,'sum us',tot)
PY
