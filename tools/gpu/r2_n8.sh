#!/bin/bash
# 8 GPUs: the default workload (config 2) under torchrun, as the driver's scaling run launches it
mkdir -p gpurun_out
N=${1:-8}; C=${2:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 \
  bench.py --gpus $N --config $C --steps 30 --warmup 5 > gpurun_out/bench_c${C}_n${N}.json 2> gpurun_out/bench_c${C}_n${N}.err
echo "rc=$?"; tail -3 gpurun_out/bench_c${C}_n${N}.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c${C}_n${N}.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','host_us_per_step','nccl','scaling')}, 'e2e', d['e2e']['value'])
PY
