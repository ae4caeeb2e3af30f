#!/bin/bash
mkdir -p gpurun_out
for k in epilogue robust full_layer chained_encoder; do
  timeout 600 python -m pytest tests/test_gpu_chain.py -q -m gpu -k $k --tb=short 2>&1 | grep -E "^(FAILED|PASSED|E   assert|[0-9]+ (passed|failed))|passed|failed|timed out|stalled" | cut -c1-250 | head -30 > gpurun_out/t_chain_$k.log
  echo "== chain $k:"; cat gpurun_out/t_chain_$k.log
done
