#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_featstore.py tests/test_gpu_dropin.py tests/test_gpu_real_configs.py \
  "tests/test_gpu_parity.py::test_pretraining_model_forward_against_oracle" \
  "tests/test_gpu_parity.py::test_output_hidden_states_against_oracle" \
  "tests/test_gpu_parity.py::test_vcr_two_head_model_and_nsp_graph_path" \
  "tests/test_gpu_train.py::test_training_trajectory_tracks_the_fp32_reference" \
  -q -m gpu --tb=short 2>&1 | tail -60 > gpurun_out/t_new.log
cat gpurun_out/t_new.log | cut -c1-300
