#!/bin/bash
# Round-2 GPU call B: whole GPU suite (no -x) on the fused sampling / LMC / Adam / k-means / sharding stack, bench lines,
# launch list and ncu --set full captures summarised ON THE BOX (the .ncu-rep files are too big to travel back).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/b_smi.txt 2>&1
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/b_pytest.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/b_pytest.log | cut -c1-220
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/b_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/b_smoke.log
echo "== bench c3"; timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_c3.json 2> gpurun_out/b_bench_c3.err; echo "rc=$?"; cut -c1-330 gpurun_out/b_bench_c3.json; tail -3 gpurun_out/b_bench_c3.err
echo "== bench default"; timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench_default.json 2> gpurun_out/b_bench_default.err; echo "rc=$?"; cut -c1-330 gpurun_out/b_bench_default.json; tail -3 gpurun_out/b_bench_default.err
echo "== bench c5 one-rank-of-8 probe (625 genes)"; timeout 900 python bench.py --config c5 --genes 625 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_bench_c5_625.json 2> gpurun_out/b_bench_c5_625.err; echo "rc=$?"; cut -c1-330 gpurun_out/b_bench_c5_625.json; tail -3 gpurun_out/b_bench_c5_625.err
echo "== ncu launch list (c3, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches_c3.csv \
  python bench.py --config c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_list.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/b_launches_c3.csv 70 > gpurun_out/b_launches_c3_summary.txt 2>&1; head -40 gpurun_out/b_launches_c3_summary.txt
echo "== ncu --set full: tcgen05 GEMM kernels of one iteration"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 7 -c 7 -o gpurun_out/b_tc \
  python bench.py --config c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_tc.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/b_tc.ncu-rep > gpurun_out/b_tc_ncu_summary.txt 2>&1; rm -f gpurun_out/b_tc.ncu-rep
echo "== ncu --set full: kernel evaluation / sampling / LL / KL / factorisation kernels"
timeout 900 ncu --set full --clock-control none -k 'regex:kmat_|sample_ll|philox|ll_fwd|ll_bwd|kl_F|kq_|pack_|potrf|trtri|gemm_dmma|feat_unpack|adam' -s 0 -c 60 -o gpurun_out/b_small \
  python bench.py --config c3 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_ncu_small.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/b_small.ncu-rep > gpurun_out/b_small_ncu_summary.txt 2>&1; rm -f gpurun_out/b_small.ncu-rep
ls -la gpurun_out | tail -20; du -sh gpurun_out
