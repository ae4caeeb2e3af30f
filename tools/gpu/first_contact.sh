#!/bin/bash
# first GPU contact: kernel unit tests in separate processes (a trap poisons the CUDA context), then parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for k in gemm attention layernorm; do
  timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k $k -x --tb=short 2>&1 | tail -60 > gpurun_out/t_$k.log
  echo "== $k: $(tail -1 gpurun_out/t_$k.log)"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short 2>&1 | tail -80 > gpurun_out/t_parity.log
echo "== parity: $(tail -1 gpurun_out/t_parity.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
