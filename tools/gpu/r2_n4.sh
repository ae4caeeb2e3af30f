#!/bin/bash
mkdir -p gpurun_out
N=4
for c in 2 4; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2975$c \
  bench.py --gpus $N --config $c --steps 30 --warmup 5 --no-extras > gpurun_out/bench_c${c}_n4.json 2> gpurun_out/bench_c${c}_n4.err
echo "config $c rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c${c}_n4.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, 'e2e', d['e2e']['value'], d.get('parity',{}).get('max_rel_err'))
PY
done
