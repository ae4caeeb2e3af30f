#!/bin/bash
# Round-2 evidence: bench line (+ reference arm), ncu launch list, ncu --set full of the dataflow chain kernel and the
# attention kernel (graphs off so every kernel is a plain launch).  CSV pages come back through gpurun_out/.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
echo "ref rc=$?"
export CPT_B200_GRAPHS=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain2_kernel -s 14 -c 2 -f -o /tmp/r02_chain \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_chain.log 2>&1
echo "ncu chain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_pp -s 14 -c 1 -f -o /tmp/r02_attn \
   python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
for n in chain attn; do ncu -i /tmp/r02_$n.ncu-rep --page raw --csv > gpurun_out/r02_${n}_raw.csv 2>/dev/null; done
ncu -i /tmp/r02_chain.ncu-rep --page source --csv > gpurun_out/r02_chain_source.csv 2>/dev/null
ncu -i /tmp/r02_chain.ncu-rep --page details --csv > gpurun_out/r02_chain_details.csv 2>/dev/null
ls -la gpurun_out/r02_*; du -sh gpurun_out
tail -3 gpurun_out/ncu_chain.log | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','host_us_per_step')}, d['e2e'], d['roofline'], d.get('parity'))
PY
