#!/bin/bash
mkdir -p gpurun_out
s=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$? $(( $(date +%s) - s )) s"
s=$(date +%s)
timeout 900 python bench.py > gpurun_out/final.json 2> gpurun_out/final.err; echo "ours rc=$? $(( $(date +%s) - s )) s"
python - <<'PY'
import json
r=json.load(open('gpurun_out/final_ref.json')); d=json.load(open('gpurun_out/final.json'))
print('reference', r['value'], r['unit'], r['cpu_baseline']['cores'], 'cores')
print({k:d.get(k) for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','higher_is_better','scaling','vs_baseline','dtype','data','gpu_launches','host_us_per_step')})
print('e2e', d['e2e']); print('roofline', d['roofline']); print('cpu_baseline', d['cpu_baseline']); print('clocks', d['clocks']); print('parity', d['parity'])
print('fresh', d['fresh_tensor_loop']['samples_per_s'], 'unmodified', d['unmodified_call']['samples_per_s'], 'other', d['other_dtype']['value'], d['other_dtype']['parity']['max_rel_err'], 'train', d['train_step']['ms_per_step'])
print('config', d['config'])
PY
