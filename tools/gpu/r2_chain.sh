#!/bin/bash
mkdir -p gpurun_out
for k in single_gemm deferred chained_encoder; do
  timeout 600 python -m pytest tests/test_gpu_chain.py -q -m gpu -k $k --tb=short 2>&1 | grep -E "^(FAILED|E   assert|E  )|passed|failed|timed out|stalled" | cut -c1-250 | head -30 > gpurun_out/t_chain_$k.log
  echo "== chain $k:"; cat gpurun_out/t_chain_$k.log
done
for st in "aod,upd,downd,qkvd" "aod" "downd" "up" "qkv"; do
  echo "##### stages $st"
  timeout 120 python tools/chain_trace.py --stages $st 2>&1 | cut -c1-300 | grep -v "timeline\|raw"
done
