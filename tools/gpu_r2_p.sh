#!/bin/bash
# Round-2 GPU call P (1 GPU): ncu evidence for the FINAL kernels -- launch lists of a C3 iteration (full shape and one
# rank of an 8-way gene split) and --set full captures of the tcgen05 kernels, summarised on the box.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== ncu launch list (c3, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p_launches_c3.csv \
  python bench.py --config c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/p_ncu_list.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/p_launches_c3.csv > gpurun_out/p_launches_c3_summary.txt 2>&1
echo "== ncu launch list (c3, 250 genes = one rank of 8)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p_launches_c3_250.csv \
  python bench.py --config c3 --genes 250 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/p_ncu_list250.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/p_launches_c3_250.csv > gpurun_out/p_launches_c3_250genes_summary.txt 2>&1
echo "== ncu --set full: tcgen05 GEMM kernels of one iteration"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 7 -c 7 -o gpurun_out/p_tc \
  python bench.py --config c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/p_ncu_tc.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/p_tc.ncu-rep > gpurun_out/p_tc_ncu_summary.txt 2>&1; rm -f gpurun_out/p_tc.ncu-rep
echo "== ncu --set full: A-bar kernel at 250 genes"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 7 -c 7 -o gpurun_out/p_tc250 \
  python bench.py --config c3 --genes 250 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/p_ncu_tc250.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/p_tc250.ncu-rep > gpurun_out/p_tc250_ncu_summary.txt 2>&1; rm -f gpurun_out/p_tc250.ncu-rep
head -12 gpurun_out/p_launches_c3_summary.txt; head -8 gpurun_out/p_launches_c3_250genes_summary.txt
rm -f gpurun_out/p_launches_c3.csv.bak; du -sh gpurun_out
