#!/bin/bash
# Round-2 GPU call R (1 GPU): final verification -- what the driver runs at round end: pytest -m gpu, smoke(), the
# default bench line (C3 headline + other_configs + cpu_baseline) and the reference arm.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "^frame #" | tail -40 > gpurun_out/r_pytest_tail.txt
tail -2 gpurun_out/r_pytest_tail.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
s=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r_bench_default.err | grep '^{' > gpurun_out/r_bench_default.json
echo "default bench rc=$? took $(( $(date +%s) - s )) s"
s=$(date +%s)
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r_bench_reference.err | grep '^{' > gpurun_out/r_bench_reference.json
echo "reference arm rc=$? took $(( $(date +%s) - s )) s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r_bench_default.json"))
print("c3", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "frac", d["roofline"]["frac"], "cpu", d["cpu_baseline"])
for o in d.get("other_configs", []):
    print(o["config"]["workload"][:30], o.get("ms_per_step"), o.get("gpu_eager_baseline"), o.get("error"))
r = json.load(open("gpurun_out/r_bench_reference.json"))
print("reference", r["value"], r["ms_per_step"], r.get("extrapolated_ms_per_step"), r["cpu_baseline"])
PY
