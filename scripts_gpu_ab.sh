#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k gemm -x --tb=short 2>&1 | tail -15 > gpurun_out/t_gemm.log
echo "== gemm: $(tail -1 gpurun_out/t_gemm.log)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short 2>&1 | tail -15 > gpurun_out/t_parity.log
echo "== parity: $(tail -1 gpurun_out/t_parity.log)"
python tools/parity_margins.py 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1]))
k=d['kernels']
print('%-22s %.3f ms/step %6.0f samples/s e2e %6.0f | '%(sys.argv[1],d['ms_per_step'],d['value'],d['e2e']['value'])+' '.join('%s %.1f'%(n.replace('gemm_',''),k[n]['us_per_launch']) for n in ('gemm_qkv','attention','gemm_attn_out','gemm_ffn_up','gemm_ffn_down','layernorm')))
PY
}
run reduce_resid X=1
run resid_in_ln CPT_B200_REDUCE_RESID=0
run reduce_resid_again X=1
