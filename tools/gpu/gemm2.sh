#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k gemm -x --tb=short 2>&1 | tail -30 > gpurun_out/t_gemm.log
echo "== gemm: $(tail -1 gpurun_out/t_gemm.log)"
timeout 600 python tools/tune_gemm.py 64 120 2>&1 | tee gpurun_out/tune_64_120.log
timeout 300 python tools/trace_gemm.py 2>&1 | tee gpurun_out/trace.log
