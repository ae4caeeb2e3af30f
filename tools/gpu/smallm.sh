#!/bin/bash
# A/B: narrower GEMM tiles for small row counts (few-shot training steps, single-query inference latency)
for v in 0 1; do
  echo "CPT_B200_SMALL_M=$v"
  for b in 4 16; do
    CPT_B200_SMALL_M=$v python tools/train_bench.py --batch $b --dropout 0.1 2>&1 | tail -1 | grep -o "ms_per_step[^,]*" | tr "\n" " "
  done
  CPT_B200_SMALL_M=$v python tools/sweep.py 2>&1 | tail -8
done
