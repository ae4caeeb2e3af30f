#!/bin/bash
mkdir -p gpurun_out
export CPT_B200_GRAPHS=0
# per-forward: 1 img gemm + 12 x (qkv, ao, up, down).  3 forwards before timed step in --profile-only (first call + warmup 16 -> too many)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 58 -c 4 -f -o gpurun_out/prof_gemm2 \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_pipe -s 13 -c 1 -f -o gpurun_out/prof_attn2 \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_attn2.log 2>&1
echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
