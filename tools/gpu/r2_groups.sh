#!/bin/bash
mkdir -p gpurun_out
for g in 1 2 3 1 2 4; do
  CPT_B200_CHAIN_GROUPS=$g timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_g$g.json 2> gpurun_out/bench_g$g.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_g$g.json'))
print('GROUPS=$g', {k:d.get(k) for k in ('value','ms_per_step')}, d.get('parity',{}).get('max_rel_err'))
PY
done
