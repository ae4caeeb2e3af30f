#!/bin/bash
# what the driver runs at round end, N GPUs: reference arm, then our arm, default flags
mkdir -p gpurun_out
N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 \
  bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/drv_ref_n$N.json 2> gpurun_out/drv_ref_n$N.err
echo "reference rc=$?"; tail -c 400 gpurun_out/drv_ref_n$N.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29732 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/drv_n$N.json 2> gpurun_out/drv_n$N.err
echo "ours rc=$?"; tail -3 gpurun_out/drv_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/drv_n$N.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches','scaling','dtype')}, d['e2e'], d.get('parity'), d['clocks'])
print(sorted(d.keys()))
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
