#!/bin/bash
# round 2: bench with the chain on and off (same box), per-kernel split
mkdir -p gpurun_out
for c in 1 0; do
  CPT_B200_CHAIN=$c timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_chain$c.json 2> gpurun_out/bench_chain$c.err
  echo "chain=$c rc=$?"; tail -2 gpurun_out/bench_chain$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_chain$c.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','model_tflops','model_frac_of_sustained_peak','clocks')})
print(d['e2e'])
for k,v in d['kernels'].items():
    print('%-16s %7.3f ms/step  %5.1f us/launch  %s'%(k,v['ms_per_step'],v['us_per_launch'],('%.0f TF'%v['tflops']) if 'tflops' in v else ''))
print('train', d.get('train_step'))
PY
done
