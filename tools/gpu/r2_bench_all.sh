#!/bin/bash
mkdir -p gpurun_out
for c in 2 4 5 3; do
  timeout 900 python bench.py --config $c --steps 30 --warmup 5 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  echo "config $c rc=$?"; tail -3 gpurun_out/bench_c$c.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_c$c.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','host_us_per_step','model_frac_of_sustained_peak','dtype')})
    print('e2e', d['e2e']['value'], 'parity', d.get('parity'), 'other', {k:v for k,v in (d.get('other_dtype') or {}).items() if k!='e2e'})
    print('fresh', d.get('fresh_tensor_loop',{}).get('samples_per_s'), d.get('fresh_tensor_loop',{}).get('graph_replays'), 'unmodified', d.get('unmodified_call',{}).get('samples_per_s'))
    print('cpu', (d.get('cpu_baseline') or {}).get('value'), 'train', (d.get('train_step') or {}).get('ms_per_step'), (d.get('train_step') or {}).get('error'))
    for k,v in (d.get('kernels') or {}).items():
        print('   %-16s %7.3f ms/step  %5.1f us/launch  %s'%(k,v['ms_per_step'],v['us_per_launch'],('%.0f TF'%v['tflops']) if 'tflops' in v else ''))
except Exception as e:
    print('parse failed', e)
PY
done
