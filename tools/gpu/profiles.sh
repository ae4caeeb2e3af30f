#!/bin/bash
# Round-1 evidence: bench line, ncu launch list, ncu --set full of the dominant kernels (graphs off so every kernel is
# a plain launch).  Reports stay in /tmp on the box; the CSV pages come back through gpurun_out/.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r01_bench_reference.json 2>> gpurun_out/r01_bench.err
echo "ref rc=$?"
export CPT_B200_GRAPHS=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 107 -c 4 -f -o /tmp/r01_gemm \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_pp -s 26 -c 1 -f -o /tmp/r01_attn \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:ln_rows -s 55 -c 1 -f -o /tmp/r01_ln \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_ln.log 2>&1
echo "ncu ln rc=$?"
for n in gemm attn ln; do ncu -i /tmp/r01_$n.ncu-rep --page raw --csv > gpurun_out/r01_${n}_raw.csv 2>/dev/null; done
ncu -i /tmp/r01_gemm.ncu-rep --page source --csv > gpurun_out/r01_gemm_source.csv 2>/dev/null
ncu -i /tmp/r01_attn.ncu-rep --page source --csv > gpurun_out/r01_attn_source.csv 2>/dev/null
ls -la gpurun_out/r01_*; du -sh gpurun_out
# training step: timing + per-class split, and the ncu launch list of one step
for args in "--batch 16" "--batch 64" "--batch 16 --T 165 --R 45" "--batch 16 --dropout 0.1"; do
  timeout 300 python tools/train_bench.py $args 2>>gpurun_out/r01_bench.err | tail -1 >> gpurun_out/r01_train_bench.jsonl
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01_train_launches.csv \
   python tools/train_bench.py --batch 64 --steps 1 --warmup 0 > gpurun_out/ncu_train_launch.log 2>&1
echo "train launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 12 -c 1 -f -o /tmp/r01_attn_bwd \
   python tools/train_bench.py --batch 64 --steps 1 --warmup 0 > gpurun_out/ncu_attn_bwd.log 2>&1
ncu -i /tmp/r01_attn_bwd.ncu-rep --page raw --csv > gpurun_out/r01_attn_bwd_raw.csv 2>/dev/null
ls -la gpurun_out/r01_*; du -sh gpurun_out
