#!/bin/bash
# Round-2 GPU call N (1 GPU): full GPU suite (spawned result managers), then the A-bar product with one epilogue group
# per accumulator stage and run-accumulated I sums: C3 headline + rank emulations.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "^frame #" | tail -60 > gpurun_out/n_pytest_tail.txt
tail -4 gpurun_out/n_pytest_tail.txt
timeout 600 python -m pytest tests/test_parallel.py -q -m gpu --tb=short 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/n_bench_c3.err | grep '^{' > gpurun_out/n_bench_c3.json
for gs in "250 8" "500 8" "1000 8"; do
  set -- $gs
  timeout 300 python bench.py --config c3 --genes $1 --samples $2 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/n_emul_$1_$2.err | grep '^{' > gpurun_out/n_emul_$1_$2.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/n_*.json")):
    try:
        d = json.load(open(f))
        print(f, "ms_per_step", round(d["ms_per_step"], 3), {k: round(v["ms_per_launch"], 3) for k, v in d["roofline"]["products"].items() if v["ms_per_launch"]}, d["clocks"].get("sm_mhz"))
    except Exception as e:
        print(f, "ERR", e)
PY
