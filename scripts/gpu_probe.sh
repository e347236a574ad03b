#!/bin/bash
# Records what the GPU box looks like (cores, memory, GPU) next to the measurements.
mkdir -p gpurun_out
{
  echo "nproc=$(nproc)"; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv
  lscpu | grep -E 'Model name|Socket|Thread|Core|MHz' | head -8
} > gpurun_out/box.txt 2>&1
cat gpurun_out/box.txt
