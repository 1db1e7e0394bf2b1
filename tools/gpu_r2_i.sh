#!/bin/bash
# Round-2 GPU call I (8 GPUs): C3 strong scaling at 2 / 4 / 8 GPUs with the final library, C4 at 8 GPUs with the
# automatic partition (Monte-Carlo samples: 500 genes would leave 62 per rank).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
port=29520
for n in 2 4 8; do
  echo "== c3 on $n GPUs"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --config c3 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/i_bench_c3_${n}gpu.err | grep '^{' > gpurun_out/i_bench_c3_${n}gpu.json
  echo "rc=$?"; cut -c1-260 gpurun_out/i_bench_c3_${n}gpu.json; tail -2 gpurun_out/i_bench_c3_${n}gpu.err | cut -c1-200
  port=$((port+1))
done
echo "== c4 on 8 GPUs (auto partition)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
  bench.py --gpus 8 --config c4 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/i_bench_c4_8gpu.err | grep '^{' > gpurun_out/i_bench_c4_8gpu.json
echo "rc=$?"; cut -c1-260 gpurun_out/i_bench_c4_8gpu.json; tail -3 gpurun_out/i_bench_c4_8gpu.err | cut -c1-300
du -sh gpurun_out
