#!/bin/bash
# 2 GPUs: the data-parallel training tests, then config 3 at N=2
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_ddp.py -q -m gpu --tb=short 2>&1 | tail -60 | cut -c1-400 > gpurun_out/t_ddp.log
cat gpurun_out/t_ddp.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus 2 --config 3 --steps 40 --warmup 5 > gpurun_out/bench_c3_2gpu.json 2> gpurun_out/bench_c3_2gpu.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_c3_2gpu.json; tail -5 gpurun_out/bench_c3_2gpu.err | cut -c1-300
