#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -40 > gpurun_out/t_all.log
cat gpurun_out/t_all.log | cut -c1-250
