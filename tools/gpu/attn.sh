#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k attention -x --tb=short 2>&1 | tail -15 > gpurun_out/t_attention.log
echo "== attention: $(tail -1 gpurun_out/t_attention.log)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short 2>&1 | tail -15 > gpurun_out/t_parity.log
echo "== parity: $(tail -1 gpurun_out/t_parity.log)"
timeout 900 python tools/sweep.py 2>&1 | tail -7
