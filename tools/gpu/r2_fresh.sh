#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -k "graph or fresh or weight_update" 2>&1 | tail -4 | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_fresh.json 2> gpurun_out/bench_fresh.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_fresh.json'))
print({k:d.get(k) for k in ('value','ms_per_step','host_us_per_step')}, 'e2e', d['e2e']['value'], 'fresh', d['fresh_tensor_loop'])
PY
