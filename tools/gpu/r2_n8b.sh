#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 \
  bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/drv_n$N.json 2> gpurun_out/drv_n$N.err
echo "ours rc=$?"; grep -v "OMP_NUM\|^\*\*\*\|NCCL version" gpurun_out/drv_n$N.err | tail -3 | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/drv_n$N.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches','host_us_per_step')}, d['e2e']['value'], d.get('parity',{}).get('max_rel_err'), d['config']['parallelism'][:60])
PY
timeout 200 python -m pytest tests/test_gpu_ddp.py -q -m gpu --tb=short -k "peer_memory" 2>&1 | tail -3 | cut -c1-300
