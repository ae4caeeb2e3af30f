#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_real_configs.py tests/test_gpu_parity.py -q -m gpu --tb=short 2>&1 | tail -8 | cut -c1-300
for e in 1 0 1 0; do
  CPT_B200_ATTN_EARLY=$e timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_early$e.json 2> gpurun_out/bench_early$e.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_early$e.json'))
print('ATTN_EARLY=$e', {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d.get('parity',{}).get('max_rel_err'))
PY
done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4
