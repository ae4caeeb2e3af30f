#!/bin/bash
# ncu --set full of the 8 backward GEMMs of the last encoder layer (wgrad / dgrad pairs of FFN-down, FFN-up,
# attention-out, QKV) in one training step at B=64, S=120
mkdir -p gpurun_out
CPT_B200_TRAIN_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 55 -c 8 -f -o /tmp/r01_bwd_gemm \
   python tools/train_bench.py --batch 64 --steps 1 --warmup 0 > gpurun_out/ncu_bwd_gemm.log 2>&1
echo "rc=$?"
ncu -i /tmp/r01_bwd_gemm.ncu-rep --page raw --csv > gpurun_out/r01_bwd_gemm_raw.csv 2>/dev/null
ls -la gpurun_out/r01_bwd_gemm_raw.csv
