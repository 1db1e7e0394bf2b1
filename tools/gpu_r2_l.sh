#!/bin/bash
# Round-2 GPU call L (1 GPU): the sample-sharded world-2 test failed once in call K -- repeat the sharding tests with
# full failure output (and once with blocking launches) to tell a race from a tolerance; C4 rank emulations.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for i in 1 2 3; do
  timeout 600 python -m pytest tests/test_parallel.py -q -m gpu --tb=short 2>&1 | grep -v Warning | tail -40 > gpurun_out/l_parallel_$i.txt
  tail -3 gpurun_out/l_parallel_$i.txt
done
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_parallel.py -q -m gpu --tb=short 2>&1 | grep -v Warning | tail -40 > gpurun_out/l_parallel_blocking.txt
tail -3 gpurun_out/l_parallel_blocking.txt
for gs in "500 1" "250 2" "125 4"; do
  set -- $gs
  timeout 300 python bench.py --config c4 --genes $1 --samples $2 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/l_emul_c4_$1_$2.err | grep '^{' > gpurun_out/l_emul_c4_$1_$2.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/l_*.json")):
    try:
        d = json.load(open(f))
        print(f, "ms_per_step", round(d["ms_per_step"], 3), {k: round(v["ms_per_launch"], 3) for k, v in d["roofline"]["products"].items() if v["ms_per_launch"]}, d["clocks"].get("sm_mhz"))
    except Exception as e:
        print(f, "ERR", e)
PY
