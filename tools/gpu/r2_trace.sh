#!/bin/bash
mkdir -p gpurun_out
for st in "ao,ln1,up,down,ln2,qkv" "ao,ln1"; do
  echo "##### stages $st"
  timeout 120 python tools/chain_trace.py --stages $st 2>&1 | cut -c1-1200
done
