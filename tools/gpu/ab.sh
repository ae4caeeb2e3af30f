#!/bin/bash
# A/B of one environment switch on the bench workload (same box, interleaved)
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  tail -2 gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1]))
k=d['kernels']['gemm_ffn_down']
print('%-22s %.3f ms/step %6.0f samples/s e2e %6.0f ffn_down %.1f us'%(sys.argv[1],d['ms_per_step'],d['value'],d['e2e']['value'],k['us_per_launch']))
PY
}
run k1 CPT_B200_DOWN_KSPLIT=1
run k2 CPT_B200_DOWN_KSPLIT=2
run k3 CPT_B200_DOWN_KSPLIT=3
run k1_again CPT_B200_DOWN_KSPLIT=1
run k4 CPT_B200_DOWN_KSPLIT=4
CPT_B200_DOWN_KSPLIT=3 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -k golden 2>&1 | tail -2
