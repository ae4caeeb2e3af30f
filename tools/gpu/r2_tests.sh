#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | tail -30 > gpurun_out/t_all.log
cat gpurun_out/t_all.log | cut -c1-300
