#!/bin/bash
# Builds the tcgen05 / fp64-pipe micro-probe (tools/mma_probe.cu) into build/mma_probe (git-ignored; travels with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Ispatial-alignment_b200/csrc -Iinclude -o build/mma_probe tools/mma_probe.cu -lcuda
echo "built build/mma_probe; run it on the GPU box: gpurun -- ./build/mma_probe"
