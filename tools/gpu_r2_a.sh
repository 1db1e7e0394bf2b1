#!/bin/bash
# Round-2 GPU call A: full GPU suite (incl. the new benchmark-shape parity cases), product timings of both forward
# forms, the C3 step with either engine, the default bench line, launch list and ncu captures.
# Every step has its own timeout and log; a failing step does not stop the rest.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/a_pytest.log
echo "== products c3"; timeout 300 python tools/bench_quadform.py --reps 5 > gpurun_out/a_products_c3.txt 2>&1; cat gpurun_out/a_products_c3.txt
echo "== products c4"; timeout 300 python tools/bench_quadform.py --M 256 --R 640000 --L 500 --reps 3 > gpurun_out/a_products_c4.txt 2>&1; cat gpurun_out/a_products_c4.txt
echo "== products c5-shaped"; timeout 300 python tools/bench_quadform.py --M 512 --R 256000 --L 160 --reps 3 > gpurun_out/a_products_c5.txt 2>&1; cat gpurun_out/a_products_c5.txt
for e in 2 1; do
  echo "== bench c3 engine $e"
  timeout 600 python bench.py --config c3 --engine $e --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_c3_e$e.json 2> gpurun_out/a_bench_c3_e$e.err
  echo "rc=$?"; cut -c1-400 gpurun_out/a_bench_c3_e$e.json
done
echo "== bench default (headline + other_configs + cpu baseline)"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err; echo "rc=$?"; cut -c1-300 gpurun_out/a_bench_default.json
echo "== bench reference arm"
timeout 1200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/a_bench_reference.json 2> gpurun_out/a_bench_reference.err; echo "rc=$?"; cut -c1-300 gpurun_out/a_bench_reference.json
echo "== ncu launch list (c3, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_launches_c3.csv \
  python bench.py --config c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu --set full: forward feature GEMM + the two backward products (1 launch each)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 8 -c 8 -o gpurun_out/a_tc \
  python bench.py --config c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_tc.log 2>&1; echo "rc=$?"
echo "== ncu --set full: kmat / sampling / LL / KL kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:kmat_|sample_|ll_fwd|ll_bwd|kl_F|kq_kernel|pack_' -s 0 -c 40 -o gpurun_out/a_small \
  python bench.py --config c3 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/a_ncu_small.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -40
