#!/bin/bash
# Round-2 GPU call H (1 GPU): verification of the final library state -- suite, smoke, product timings, C3 bench line.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/h_pytest.log | cut -c1-200
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/h_smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/h_smoke.log
echo "== products c3"; timeout 300 python tools/bench_quadform.py --reps 5 2>&1 | tee gpurun_out/h_products_c3.txt
echo "== bench c3"; timeout 600 python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench_c3.json 2> gpurun_out/h_bench_c3.err; echo "rc=$?"; cut -c1-300 gpurun_out/h_bench_c3.json; tail -3 gpurun_out/h_bench_c3.err
du -sh gpurun_out
