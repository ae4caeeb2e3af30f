#!/bin/bash
mkdir -p gpurun_out
export CPT_B200_CHAIN_MIN_ROWS=256
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python -m pytest \
  "tests/test_gpu_chain.py::test_layer_chain_with_deferred_layernorm" -q -m gpu -x --tb=short -k "1000" > gpurun_out/r02_racecheck_chain.log 2>&1
echo "racecheck rc=$?"
tail -25 gpurun_out/r02_racecheck_chain.log | cut -c1-300
