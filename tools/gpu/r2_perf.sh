#!/bin/bash
mkdir -p gpurun_out
for g in 1 2 3 4; do
for st in "ao,ln1,up,down,ln2,qkv" "aoln,up,downln,qkv"; do
  echo "##### groups $g stages $st"
  CPT_B200_CHAIN_GROUPS=$g timeout 120 python tools/chain_trace.py --stages $st 2>&1 | cut -c1-200 | grep -v "timeline\|raw"
done
done
