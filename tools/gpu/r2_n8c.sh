#!/bin/bash
mkdir -p gpurun_out
N=8
for c in 4 5; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2976$c \
  bench.py --gpus $N --config $c --steps 20 --warmup 4 --no-extras > gpurun_out/bench_c${c}_n8.json 2> gpurun_out/bench_c${c}_n8.err
echo "config $c rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_c${c}_n8.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, 'e2e', d['e2e']['value'], d.get('parity',{}).get('max_rel_err'))
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_c${c}_n8.err').read()[-600:])
PY
done
