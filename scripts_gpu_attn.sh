#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k attention -x --tb=short 2>&1 | tail -15 > gpurun_out/t_attention.log
echo "== attention: $(tail -1 gpurun_out/t_attention.log)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short 2>&1 | tail -15 > gpurun_out/t_parity.log
echo "== parity: $(tail -1 gpurun_out/t_parity.log)"
timeout 300 python tools/trace_attn.py 2>&1 | tail -11
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
k=d['kernels']
print('%.3f ms/step %6.0f samples/s e2e %6.0f | '%(d['ms_per_step'],d['value'],d['e2e']['value'])+' '.join('%s %.1f'%(n.replace('gemm_',''),k[n]['us_per_launch']) for n in ('gemm_qkv','attention','gemm_attn_out','gemm_ffn_up','gemm_ffn_down','layernorm')))
PY
for cfg in "gemm_qkv:192:2" "gemm_qkv:192:1" "gemm_qkv:256:1"; do
CPT_B200_GEMM=$cfg timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python - "$cfg" <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_x.json'))
k=d['kernels']
print('%-18s %.3f ms/step %6.0f samples/s | '%(sys.argv[1],d['ms_per_step'],d['value'])+' '.join('%s %.1f'%(n.replace('gemm_',''),k[n]['us_per_launch']) for n in ('gemm_qkv','attention','gemm_attn_out','gemm_ffn_up','gemm_ffn_down','layernorm')))
PY
done
