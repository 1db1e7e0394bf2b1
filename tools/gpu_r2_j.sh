#!/bin/bash
# Round-2 GPU call J (1 GPU): hybrid (genes x samples) sharding tests at world 4 on one device, then the per-rank
# compute of every 8-rank grid of C3 emulated on one GPU (--genes P/Wg --samples S/Ws): picks the automatic grid.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_parallel.py -q -m gpu 2>&1 | tail -5
for gs in "250 8" "500 4" "1000 2" "2000 1" "500 8" "1000 4" "2000 2" "1000 8" "2000 4"; do
  set -- $gs
  echo "== c3 rank emulation: genes=$1 samples=$2"
  timeout 300 python bench.py --config c3 --genes $1 --samples $2 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/j_emul_$1_$2.err | grep '^{' > gpurun_out/j_emul_$1_$2.json
  python - <<PY
import json
d=json.load(open("gpurun_out/j_emul_$1_$2.json"))
print("ms_per_step", round(d["ms_per_step"],3), {k: round(v["ms_per_launch"],3) for k,v in d["roofline"]["products"].items() if v["ms_per_launch"]})
PY
done
du -sh gpurun_out
