#!/bin/bash
# round 2: the dataflow chain kernel — unit tests, encoder parity, event log, bench with the chain on / off
mkdir -p gpurun_out
for k in single_gemm accumulate epilogue robust full_layer chained_encoder; do
  timeout 600 python -m pytest tests/test_gpu_chain.py -q -m gpu -k $k -x --tb=short 2>&1 | tail -25 > gpurun_out/t_chain_$k.log
  echo "== chain $k: $(tail -1 gpurun_out/t_chain_$k.log)"
done
for st in "aoln,up,downln,qkv" "aoln" "downln" "aoln,up"; do
  echo "##### stages $st"
  timeout 120 python tools/chain_trace.py --stages $st 2>&1 | cut -c1-700
done
for c in 1 0; do
  CPT_B200_CHAIN=$c timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_chain$c.json 2> gpurun_out/bench_chain$c.err
  echo "chain=$c rc=$?"; tail -2 gpurun_out/bench_chain$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_chain$c.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','model_frac_of_sustained_peak','clocks')})
for k,v in d['kernels'].items():
    print('%-16s %7.3f ms/step  %5.1f us/launch  %s'%(k,v['ms_per_step'],v['us_per_launch'],('%.0f TF'%v['tflops']) if 'tflops' in v else ''))
PY
done
