#!/bin/bash
# Round-2 GPU call D: suite on the torch custom-op layer + lockstep producers; product timings; bench lines.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/d_pytest.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/d_pytest.log | cut -c1-200
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/d_smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/d_smoke.log
echo "== products c3"; timeout 300 python tools/bench_quadform.py --reps 5 2>&1 | tee gpurun_out/d_products_c3.txt
echo "== products c4"; timeout 300 python tools/bench_quadform.py --M 256 --R 640000 --L 500 --reps 3 2>&1 | tee gpurun_out/d_products_c4.txt
echo "== products c5 rank shape"; timeout 600 python tools/bench_quadform.py --M 512 --R 6400000 --L 625 --reps 2 2>&1 | tee gpurun_out/d_products_c5.txt
echo "== bench c3"; timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench_c3.json 2> gpurun_out/d_bench_c3.err; echo "rc=$?"; cut -c1-330 gpurun_out/d_bench_c3.json; tail -3 gpurun_out/d_bench_c3.err
for c in c1 c2; do
  echo "== bench $c eager (no graph)"; timeout 300 python bench.py --config $c --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/d_bench_${c}_eager.json 2> gpurun_out/d_bench_${c}_eager.err; echo "rc=$?"; cut -c1-250 gpurun_out/d_bench_${c}_eager.json
  echo "== bench $c graph"; timeout 300 python bench.py --config $c --graph --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/d_bench_${c}_graph.json 2> gpurun_out/d_bench_${c}_graph.err; echo "rc=$?"; cut -c1-250 gpurun_out/d_bench_${c}_graph.json
done
du -sh gpurun_out
