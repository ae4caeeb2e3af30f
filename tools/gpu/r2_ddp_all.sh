#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_ddp.py -q -m gpu --tb=short 2>&1 | tail -12 | cut -c1-300 > gpurun_out/t_ddp.log
cat gpurun_out/t_ddp.log
