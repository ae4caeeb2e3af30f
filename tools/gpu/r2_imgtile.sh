#!/bin/bash
mkdir -p gpurun_out
for g in "" "gemm_img:128:1" "gemm_img:128:2" "gemm_img:192:1" "gemm_img:64:1"; do
  CPT_B200_GEMM="$g" timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_t.json'))
    k=d['kernels']
    print('GEMM="$g"', d['ms_per_step'], 'gemm_img', k['gemm_img']['us_per_launch'], 'qkv0', k['gemm_qkv']['us_per_launch'], d['parity']['max_rel_err'])
except Exception as e:
    print('GEMM="$g" failed', e, open('gpurun_out/bench_t.err').read()[-300:])
PY
done
