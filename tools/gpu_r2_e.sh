#!/bin/bash
# Round-2 GPU call E: K-block-major G^T for the Omega-bar product -- suite + product timings at C3 / C4 / C5-rank shapes.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/e_pytest.log | cut -c1-200
echo "== products c3"; timeout 300 python tools/bench_quadform.py --reps 5 2>&1 | tee gpurun_out/e_products_c3.txt
echo "== products c4"; timeout 300 python tools/bench_quadform.py --M 256 --R 640000 --L 500 --reps 3 2>&1 | tee gpurun_out/e_products_c4.txt
echo "== products c5 rank shape"; timeout 600 python tools/bench_quadform.py --M 512 --R 6400000 --L 625 --reps 2 2>&1 | tee gpurun_out/e_products_c5.txt
echo "== products c3 rank-of-8 shape (250 genes)"; timeout 300 python tools/bench_quadform.py --L 250 --reps 5 2>&1 | tee gpurun_out/e_products_c3_250.txt
du -sh gpurun_out
