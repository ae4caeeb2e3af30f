#!/bin/bash
# needs >= 2 GPUs: NCCL paths (DDP tests, logits all-gather, in-backward gradient all-reduce) and the N = 2 bench lines
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_ddp.py -q -m gpu --tb=short 2>&1 | tail -8 | cut -c1-250 | tee gpurun_out/t_ddp.log
for c in 2 4 3; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c --steps 30 --warmup 5 > gpurun_out/bench_c${c}_n2.json 2> gpurun_out/bench_c${c}_n2.err
  echo "config $c N=2 rc=$?"; tail -3 gpurun_out/bench_c${c}_n2.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_c${c}_n2.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','host_us_per_step','nccl')}, 'e2e', d['e2e']['value'])
except Exception as e:
    print('parse failed', e)
PY
done
timeout 600 python bench.py --config 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 same box', d['value'], d['ms_per_step'])"
