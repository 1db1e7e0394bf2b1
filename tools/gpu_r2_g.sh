#!/bin/bash
# Round-2 GPU call G (1 GPU): final-state evidence -- suite, default bench line, reference arm, launch lists (C3 and one
# rank of 8), ncu --set full summaries (summarised on the box: the reports are too big to travel).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/g_smi.txt 2>&1
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/g_pytest.log | cut -c1-200
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/g_smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/g_smoke.log
echo "== bench default (driver's command line)"
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/g_bench_default.json 2> gpurun_out/g_bench_default.err; echo "rc=$?"; cut -c1-300 gpurun_out/g_bench_default.json; tail -3 gpurun_out/g_bench_default.err
echo "== bench reference arm"
timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/g_bench_reference.json 2> gpurun_out/g_bench_reference.err; echo "rc=$?"; cut -c1-300 gpurun_out/g_bench_reference.json
echo "== bench one rank of 8 (c3, 250 genes)"
timeout 600 python bench.py --config c3 --genes 250 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g_bench_c3_250.json 2> gpurun_out/g_bench_c3_250.err; echo "rc=$?"; cut -c1-200 gpurun_out/g_bench_c3_250.json
echo "== ncu launch list (c3, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_launches_c3.csv \
  python bench.py --config c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g_ncu_list.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/g_launches_c3.csv 80 > gpurun_out/g_launches_c3_summary.txt 2>&1; head -30 gpurun_out/g_launches_c3_summary.txt; rm -f gpurun_out/g_launches_c3.csv
echo "== ncu launch list (c3, 250 genes = one rank of 8)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_launches_c3_250.csv \
  python bench.py --config c3 --genes 250 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g_ncu_list250.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/g_launches_c3_250.csv 80 > gpurun_out/g_launches_c3_250_summary.txt 2>&1; head -45 gpurun_out/g_launches_c3_250_summary.txt; rm -f gpurun_out/g_launches_c3_250.csv
echo "== ncu --set full: tcgen05 GEMM kernels of one iteration"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 7 -c 7 -o gpurun_out/g_tc \
  python bench.py --config c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/g_ncu_tc.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/g_tc.ncu-rep > gpurun_out/g_tc_ncu_summary.txt 2>&1; rm -f gpurun_out/g_tc.ncu-rep
echo "== ncu --set full: the other kernels of one iteration"
timeout 900 ncu --set full --clock-control none -k 'regex:kmat_|sample_ll|philox|ll_fwd|ll_bwd|kl_F|kl_G|kq_|pack_|potrf|trtri|gemm_dmma|gemm_strided|feat_unpack|adam|mirror|warp_predict' -s 0 -c 90 -o gpurun_out/g_small \
  python bench.py --config c3 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/g_ncu_small.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/g_small.ncu-rep > gpurun_out/g_small_ncu_summary.txt 2>&1; rm -f gpurun_out/g_small.ncu-rep
du -sh gpurun_out
