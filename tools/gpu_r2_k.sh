#!/bin/bash
# Round-2 GPU call K (1 GPU): full GPU suite after (1) two epilogue groups + a-prefetch in the A-bar product, (2) the
# Omega_F chain as its own autograd node on a side stream, (3) K_uu of the data layer factorised ahead of the layer;
# then the C3 headline and the per-rank emulations of an 8- and 4-way split.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/k_pytest_tail.txt
tail -8 gpurun_out/k_pytest_tail.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/k_bench_c3.err | grep '^{' > gpurun_out/k_bench_c3.json
for gs in "250 8" "500 4" "500 8" "2000 1"; do
  set -- $gs
  timeout 300 python bench.py --config c3 --genes $1 --samples $2 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/k_emul_$1_$2.err | grep '^{' > gpurun_out/k_emul_$1_$2.json
done
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/k_bench_c4.err | grep '^{' > gpurun_out/k_bench_c4.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/k_*.json")):
    try:
        d = json.load(open(f))
        print(f, "ms_per_step", round(d["ms_per_step"], 3), {k: round(v["ms_per_launch"], 3) for k, v in d["roofline"]["products"].items() if v["ms_per_launch"]}, d["clocks"].get("sm_mhz"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/k_bench_c3.err
du -sh gpurun_out
