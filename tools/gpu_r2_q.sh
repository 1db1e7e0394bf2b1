#!/bin/bash
# Round-2 GPU call Q (N GPUs, N = $1): C3 gene-sharded on N GPUs with the final library (scaling table of DESIGN.md 7).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
n=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2967$n \
  bench.py --gpus $n --config c3 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/q_bench_c3_${n}gpu.err | grep '^{' > gpurun_out/q_bench_c3_${n}gpu.json
echo "rc=$?"; cut -c1-200 gpurun_out/q_bench_c3_${n}gpu.json; tail -2 gpurun_out/q_bench_c3_${n}gpu.err | cut -c1-200
