#!/bin/bash
# Round-2 GPU call O (8 GPUs): C3 on 8 GPUs with the final library -- gene sharding (the automatic partition) and the
# 4 x 2 genes x samples grid (HybridSharding over NCCL sub-groups).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
port=29620
echo "== c3 on 8 GPUs, genes"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
  bench.py --gpus 8 --config c3 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/o_bench_c3_8gpu_genes.err | grep '^{' > gpurun_out/o_bench_c3_8gpu_genes.json
echo "rc=$?"; cut -c1-200 gpurun_out/o_bench_c3_8gpu_genes.json; tail -2 gpurun_out/o_bench_c3_8gpu_genes.err | cut -c1-200
port=$((port+1))
echo "== c3 on 8 GPUs, 4 gene slices x 2 sample groups"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
  bench.py --gpus 8 --config c3 --sample-groups 2 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/o_bench_c3_8gpu_4x2.err | grep '^{' > gpurun_out/o_bench_c3_8gpu_4x2.json
echo "rc=$?"; cut -c1-200 gpurun_out/o_bench_c3_8gpu_4x2.json; tail -2 gpurun_out/o_bench_c3_8gpu_4x2.err | cut -c1-200
python - <<'PY'
import json
for f in ("genes", "4x2"):
    try:
        d = json.load(open(f"gpurun_out/o_bench_c3_8gpu_{f}.json"))
        print(f, d["ms_per_step"], d["value"], d["loss_last"] if "loss_last" in d else d.get("config"))
    except Exception as e:
        print(f, "ERR", e)
PY
