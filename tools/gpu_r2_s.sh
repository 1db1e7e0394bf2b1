#!/bin/bash
# Round-2 GPU call S (1 GPU): verification of the final library rebuilt from HEAD -- what the driver runs at round end:
# pytest -m gpu, smoke(), the default bench line (C3 headline + other_configs + cpu_baseline), then the reference arm.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
s=$(date +%s)
timeout 420 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | grep -v "^frame #" | tail -40 > gpurun_out/s_pytest_tail.txt
echo "pytest rc=${PIPESTATUS[0]} took $(( $(date +%s) - s )) s"; tail -2 gpurun_out/s_pytest_tail.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
s=$(date +%s)
timeout 420 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/s_bench_default.err | grep '^{' > gpurun_out/s_bench_default.json
echo "default bench rc=${PIPESTATUS[0]} took $(( $(date +%s) - s )) s"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/s_bench_default.json"))
    print("c3", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "frac", d["roofline"]["frac"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"])
    for o in d.get("other_configs", []):
        print(o["config"]["workload"][:30], o.get("ms_per_step"), o.get("gpu_eager_baseline"), o.get("error"))
except Exception as e:
    print("bench line unreadable:", e)
PY
tail -3 gpurun_out/s_bench_default.err | cut -c1-300
s=$(date +%s)
timeout 240 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/s_bench_reference.err | grep '^{' > gpurun_out/s_bench_reference.json
echo "reference arm rc=${PIPESTATUS[0]} took $(( $(date +%s) - s )) s"; cut -c1-400 gpurun_out/s_bench_reference.json
