#!/bin/bash
# Next-round checklist for the experimental cta_group::2 forward kernel (run on the GPU box through gpurun):
#   1. its opt-in unit cases (even / odd row-tile counts, few / many pair items),
#   2. the whole GPU suite with the kernel selected,
#   3. the forward product alone and the C3 step with and without it.
# If all of it is green and (3) shows a gain, make it the default in gpsa_quadform_fwd_tc (tc_quadform.cu: want_pair).
set -x
mkdir -p gpurun_out
GPSA_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -k pair 2>&1 | tail -3
# the opt-in C3-sized property test, default kernel and pair kernel (make it unconditional once it has passed here)
GPSA_TEST_FULLSIZE=1 timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -k full_size 2>&1 | tail -3
GPSA_FWD_PAIR=1 GPSA_TEST_FULLSIZE=1 timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -k full_size 2>&1 | tail -3
GPSA_FWD_PAIR=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
  GPSA_FWD_PAIR=$v timeout 120 python tools/bench_quadform.py --which fwd --reps 3 2>&1 | tail -1
  GPSA_FWD_PAIR=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-220
done | tee gpurun_out/verify_pair.txt
