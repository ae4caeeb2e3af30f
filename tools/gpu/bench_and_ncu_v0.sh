#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 58 -c 4 -f -o gpurun_out/prof_gemm \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 13 -c 1 -f -o gpurun_out/prof_attn \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
ls -la gpurun_out
