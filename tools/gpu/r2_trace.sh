#!/bin/bash
mkdir -p gpurun_out
for st in "ln1" "ln1,up"; do
  echo "##### stages $st"
  timeout 120 python tools/chain_trace.py --stages $st 2>&1 | cut -c1-400
done
