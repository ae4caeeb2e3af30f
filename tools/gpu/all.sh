#!/bin/bash
mkdir -p gpurun_out
for k in gemm attention layernorm; do
  timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k $k -x --tb=short 2>&1 | tail -40 > gpurun_out/t_$k.log
  echo "== $k: $(tail -1 gpurun_out/t_$k.log)"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short 2>&1 | tail -80 > gpurun_out/t_parity.log
echo "== parity: $(tail -1 gpurun_out/t_parity.log)"
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','model_tflops','model_frac_of_sustained_peak','clocks')})
print(d['e2e'], d.get('cpu_baseline',{}).get('value'))
for k,v in d['kernels'].items():
    print('%-16s %7.3f ms/step  %5.1f us/launch  %s'%(k,v['ms_per_step'],v['us_per_launch'],('%.0f TF'%v['tflops']) if 'tflops' in v else ''))
PY
