#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
run() {  # config, collective
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 \
    bench.py --gpus $N --config $1 --collective $2 --steps 30 --warmup 5 --no-extras > gpurun_out/bench_c$1_n${N}_$2.json 2> gpurun_out/bench_c$1_n${N}_$2.err
  echo "config $1 $2 rc=$?"; tail -2 gpurun_out/bench_c$1_n${N}_$2.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_c$1_n${N}_$2.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','host_us_per_step')}, 'e2e', d['e2e']['value'], d.get('parity'))
except Exception as e:
    print('parse failed', e)
PY
}
run 2 peer; run 2 nccl; run 4 peer
