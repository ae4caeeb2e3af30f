#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k gemm -x --tb=short 2>&1 | tail -5 > gpurun_out/t_gemm.log
echo "== gemm: $(tail -1 gpurun_out/t_gemm.log)"
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','model_tflops','model_frac_of_sustained_peak','clocks')})
print(d['e2e'])
for k,v in d['kernels'].items():
    print('%-16s %7.3f ms/step  %5.1f us/launch  %s'%(k,v['ms_per_step'],v['us_per_launch'],('%.0f TF'%v['tflops']) if 'tflops' in v else ''))
PY
