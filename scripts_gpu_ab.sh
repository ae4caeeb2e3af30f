#!/bin/bash
mkdir -p gpurun_out
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
run default X=1
run down192p CPT_B200_GEMM=gemm_ffn_down:192:2
run ao192p CPT_B200_GEMM=gemm_attn_out:192:2
run down128p_ao128p CPT_B200_GEMM=gemm_ffn_down:128:2,gemm_attn_out:128:2
run ao64 CPT_B200_GEMM=gemm_attn_out:64:1
run qkv192p_up192p CPT_B200_GEMM=gemm_qkv:192:2,gemm_ffn_up:192:2
