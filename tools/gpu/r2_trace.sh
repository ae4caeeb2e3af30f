#!/bin/bash
mkdir -p gpurun_out
for st in "aod,upd,downd,qkvd" "upd" "qkvd"; do
  echo "##### stages $st"
  timeout 120 python tools/chain_trace.py --stages $st 2>&1 | grep -v "^LN task\|^fused tile\|^pair" | cut -c1-330
done
