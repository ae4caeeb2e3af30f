#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  tail -2 gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1]))
print('%-22s %.3f ms/step %6.0f samples/s e2e %6.0f launches %d'%(sys.argv[1],d['ms_per_step'],d['value'],d['e2e']['value'],d['gpu_launches']))
PY
}
run split1 CPT_B200_SPLIT=1
run split2 CPT_B200_SPLIT=2
run split1_again CPT_B200_SPLIT=1
CPT_B200_SPLIT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short 2>&1 | tail -3
