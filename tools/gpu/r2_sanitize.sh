#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels at small shapes (chain kernels, fetch warp, scoring, input assembly)
mkdir -p gpurun_out
export CPT_B200_CHAIN_MIN_ROWS=256
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest \
  "tests/test_gpu_chain.py::test_single_gemm_stage_16bit_out" \
  "tests/test_gpu_chain.py::test_layer_chain_with_deferred_layernorm" \
  "tests/test_gpu_chain.py::test_chained_encoder_matches_unfused_and_oracle" \
  tests/test_gpu_scoring.py tests/test_featstore.py \
  "tests/test_gpu_parity.py::test_multi_mask_rows_for_the_visual_genome_caller" \
  -q -m gpu -x --tb=short -k "not 7680 and not 3000" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"
tail -15 gpurun_out/r02_sanitizer_memcheck.log | cut -c1-300
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r02_sanitizer_memcheck.log
