#!/bin/bash
# round 2: the dataflow chain kernel — micro-experiment, unit tests, encoder parity, bench with the chain on / off
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 120 tools/micro/mix_cta_group > gpurun_out/mix_cta_group.log 2>&1; echo "mix rc=$?"; cat gpurun_out/mix_cta_group.log
for k in single_gemm accumulate full_layer chained_encoder; do
  timeout 600 python -m pytest tests/test_gpu_chain.py -q -m gpu -k $k -x --tb=short 2>&1 | tail -40 > gpurun_out/t_chain_$k.log
  echo "== chain $k: $(tail -1 gpurun_out/t_chain_$k.log)"
done
