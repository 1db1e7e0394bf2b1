#!/bin/bash
# Round-2 GPU call M (1 GPU): catch the intermittent exception of the multi-process sharding tests with its message.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for i in 1 2 3 4 5 6; do
  timeout 600 python -m pytest tests/test_parallel.py -q -m gpu --tb=long > gpurun_out/m_parallel_$i.log 2>&1
  tail -1 gpurun_out/m_parallel_$i.log
  if grep -q "failed" gpurun_out/m_parallel_$i.log; then
    grep -n "Error\|error\|raise\|Exception" gpurun_out/m_parallel_$i.log | grep -v "^.*frame #" | head -40 > gpurun_out/m_fail_$i.txt
    grep -v "^frame #" gpurun_out/m_parallel_$i.log | head -150 >> gpurun_out/m_fail_$i.txt
    rm gpurun_out/m_parallel_$i.log
    break
  fi
  rm gpurun_out/m_parallel_$i.log
done
ls gpurun_out | grep m_
