#!/bin/bash
# Round-2 GPU call U (1 GPU): verification of the final library (clean-up build) -- what the driver runs at round end:
# pytest -m gpu, smoke(), the default bench line; then the ncu launch list of a C3 iteration and the reference arm.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
s=$(date +%s)
timeout 300 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | grep -v "^frame #" | tail -40 > gpurun_out/u_pytest_tail.txt
echo "pytest rc=${PIPESTATUS[0]} took $(( $(date +%s) - s )) s"; tail -2 gpurun_out/u_pytest_tail.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
s=$(date +%s)
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/u_bench_default.err | grep '^{' > gpurun_out/u_bench_default.json
echo "default bench rc=${PIPESTATUS[0]} took $(( $(date +%s) - s )) s"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/u_bench_default.json"))
    print("c3", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "frac", d["roofline"]["frac"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"])
    for o in d.get("other_configs", []):
        print(o["config"]["workload"][:30], o.get("ms_per_step"), o.get("gpu_eager_baseline"), o.get("error"))
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/u_launches_c3.csv \
  python bench.py --config c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/u_ncu_list.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/u_launches_c3.csv 60 > gpurun_out/u_launches_c3_summary.txt 2>&1
head -12 gpurun_out/u_launches_c3_summary.txt
timeout 100 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/u_bench_reference.err | grep '^{' > gpurun_out/u_bench_reference.json
echo "reference arm rc=${PIPESTATUS[0]}"; cut -c1-300 gpurun_out/u_bench_reference.json
