#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ddp.py -q -m gpu --tb=short -k "peer_memory" 2>&1 | tail -30 | cut -c1-400
