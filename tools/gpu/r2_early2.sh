#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_real_configs.py tests/test_gpu_parity.py tests/test_gpu_kernels.py -q -m gpu --tb=short -x 2>&1 | tail -8 | cut -c1-300
for e in 1 0 1 0; do
  CPT_B200_CHAIN_EARLY=$e timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_cearly$e.json 2> gpurun_out/bench_cearly$e.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_cearly$e.json'))
    print('CHAIN_EARLY=$e', {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d.get('parity',{}).get('max_rel_err'))
except Exception as ex:
    print('CHAIN_EARLY=$e failed', ex); print(open('gpurun_out/bench_cearly$e.err').read()[-800:])
PY
done
for c in 4 5; do
for e in 1 0; do
  CPT_B200_CHAIN_EARLY=$e timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_cearly.json 2> gpurun_out/bench_cearly.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_cearly.json'))
print('config $c CHAIN_EARLY=$e', {k:d.get(k) for k in ('value','ms_per_step')}, d.get('parity',{}).get('max_rel_err'))
PY
done; done
