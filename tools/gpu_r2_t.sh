#!/bin/bash
# Round-2 GPU call T (1 GPU): the HBM-bound / small-GEMM clean-ups (64-wide transposing packs, tiled feat_unpack and
# mirror_lower, 104 x 104 fp32 tiles for the M = 200 algebra, 32 x 32 fp64 tiles for single M x M products, no
# materialised zero gradients) -- full GPU suite, then A/B against the previous library (tools/ab/, same box).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
LIB=spatial-alignment_b200/gpsa/libgpsa_b200.so
s=$(date +%s)
timeout 420 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | grep -v "^frame #" | tail -40 > gpurun_out/t_pytest_tail.txt
echo "pytest rc=${PIPESTATUS[0]} took $(( $(date +%s) - s )) s"; tail -2 gpurun_out/t_pytest_tail.txt
b() { timeout 200 python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline "${@:2}" 2> gpurun_out/t_$1.err | grep '^{' > gpurun_out/t_$1.json
  python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/t_{sys.argv[1]}.json"))
    p = d["roofline"]["products"]
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "products %.2f %.2f %.2f" % tuple(p[k]["ms_per_launch"] for k in ("fwd", "bwd_alpha", "bwd_omega")), "sm_mhz", d["clocks"]["sm_mhz"], "loss", d["loss_last"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
b new_c3; b new_250 --genes 250
cp $LIB /tmp/new.so; cp tools/ab/libgpsa_b200_old.so $LIB
b old_c3; b old_250 --genes 250
cp /tmp/new.so $LIB
b new_c3_again
echo "== ncu launch list (new library, c3, 2 steps)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t_launches_c3.csv \
  python bench.py --config c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/t_ncu_list.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/t_launches_c3.csv 60 > gpurun_out/t_launches_c3_summary.txt 2>&1
head -50 gpurun_out/t_launches_c3_summary.txt
