#!/bin/bash
mkdir -p gpurun_out
for g in 2 3 5 10; do
  echo "##### groups $g"
  CPT_B200_CHAIN_GROUPS=$g timeout 120 python tools/chain_trace.py --stages aod,upd,downd,qkvd 2>&1 | cut -c1-200 | grep -v "timeline\|raw"
done
for c in 1 0; do
  CPT_B200_CHAIN=$c timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_chain$c.json 2> gpurun_out/bench_chain$c.err
  echo "chain=$c rc=$?"; tail -2 gpurun_out/bench_chain$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_chain$c.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','model_frac_of_sustained_peak','clocks')})
print(d['e2e'])
for k,v in d['kernels'].items():
    print('%-16s %7.3f ms/step  %5.1f us/launch  %s'%(k,v['ms_per_step'],v['us_per_launch'],('%.0f TF'%v['tflops']) if 'tflops' in v else ''))
PY
done
