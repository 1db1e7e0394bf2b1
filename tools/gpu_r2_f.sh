#!/bin/bash
# Round-2 GPU call F (8 GPUs): the full C5 configuration sharded over 8 B200, and C3 at 8 GPUs.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/f_smi.txt 2>&1
echo "== c5 on 8 GPUs"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --config c5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_bench_c5_8gpu.json 2> gpurun_out/f_bench_c5_8gpu.err
echo "rc=$?"; tail -c 3000 gpurun_out/f_bench_c5_8gpu.json | cut -c1-1500; tail -5 gpurun_out/f_bench_c5_8gpu.err | cut -c1-300
echo "== c3 on 8 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 8 --config c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_c3_8gpu.json 2> gpurun_out/f_bench_c3_8gpu.err
echo "rc=$?"; tail -c 3000 gpurun_out/f_bench_c3_8gpu.json | cut -c1-600; tail -3 gpurun_out/f_bench_c3_8gpu.err | cut -c1-300
echo "== c4 on 8 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 8 --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_c4_8gpu.json 2> gpurun_out/f_bench_c4_8gpu.err
echo "rc=$?"; tail -c 3000 gpurun_out/f_bench_c4_8gpu.json | cut -c1-600; tail -3 gpurun_out/f_bench_c4_8gpu.err | cut -c1-300
du -sh gpurun_out
