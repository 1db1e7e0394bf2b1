#!/bin/bash
# Round-2 GPU call C: whole GPU suite after the chained-accumulation rewrite of the tcgen05 GEMM core, the fp32 Omega_F
# path, the callable slow path; product timings at the C3 / C4 / C5-rank shapes; bench lines; launch list.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/c_pytest.log | cut -c1-200
echo "== products c3"; timeout 300 python tools/bench_quadform.py --reps 5 2>&1 | tee gpurun_out/c_products_c3.txt
echo "== products c4"; timeout 300 python tools/bench_quadform.py --M 256 --R 640000 --L 500 --reps 3 2>&1 | tee gpurun_out/c_products_c4.txt
for R in 800000 6400000; do
  echo "== products c5 rank shape R=$R"; timeout 600 python tools/bench_quadform.py --M 512 --R $R --L 625 --reps 2 2>&1 | tee gpurun_out/c_products_c5_$R.txt
done
echo "== bench c3"; timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_c3.json 2> gpurun_out/c_bench_c3.err; echo "rc=$?"; cut -c1-330 gpurun_out/c_bench_c3.json; tail -3 gpurun_out/c_bench_c3.err
echo "== bench c4"; timeout 600 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_c4.json 2> gpurun_out/c_bench_c4.err; echo "rc=$?"; cut -c1-330 gpurun_out/c_bench_c4.json; tail -3 gpurun_out/c_bench_c4.err
echo "== ncu launch list (c3, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_launches_c3.csv \
  python bench.py --config c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu_list.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/c_launches_c3.csv 70 > gpurun_out/c_launches_c3_summary.txt 2>&1; head -45 gpurun_out/c_launches_c3_summary.txt
rm -f gpurun_out/c_launches_c3.csv
du -sh gpurun_out
