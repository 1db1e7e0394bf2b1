#!/bin/bash
# GPU bring-up ladder for the tcgen05 engine: each rung in its own process (a trap poisons the CUDA context),
# each under a hard timeout so a protocol bug cannot hang the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/tc_ladder.log 2>&1
for t in "test_gemm_core" "test_gemm_tc_generic" "test_quadform_tc_fwd_bwd" "test_tc_unsupported_M" "test_data_layer_engines_agree"; do
  echo "=== $t" >> gpurun_out/tc_ladder.log
  timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu -k "$t" --timeout 120 -x 2>&1 | tail -40 >> gpurun_out/tc_ladder.log
done
tail -120 gpurun_out/tc_ladder.log
