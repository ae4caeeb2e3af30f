#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_real_configs.py -q -m gpu --tb=short 2>&1 | tail -15 | cut -c1-300
for st in "aod,upd,downd,qkvd" "upd"; do
  timeout 120 python tools/chain_trace.py --stages $st 2>&1 | grep -v "^LN task\|^fused tile\|^pair [37]" | head -12 | cut -c1-330
done
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','host_us_per_step')}, d['e2e']['value'], d.get('parity'), d['roofline']['us_per_launch'])
PY
