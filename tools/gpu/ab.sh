#!/bin/bash
# A/B of one environment switch on the bench workload (same box, interleaved)
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1]))
k=d['kernels']['layernorm']
print('%-22s %.3f ms/step %6.0f samples/s e2e %6.0f layernorm %.1f us'%(sys.argv[1],d['ms_per_step'],d['value'],d['e2e']['value'],k['us_per_launch']))
PY
}
run ln2 CPT_B200_LN_CTAS=2
run ln3 CPT_B200_LN_CTAS=3
run ln4 CPT_B200_LN_CTAS=4
run ln2_again CPT_B200_LN_CTAS=2
run ln6 CPT_B200_LN_CTAS=6
