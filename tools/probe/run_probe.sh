#!/bin/bash
# Runs every probe variant in its own process (a watchdog trap poisons the CUDA context).
cd "$(dirname "$0")"
mkdir -p ../../gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader
for shape in "208 208" "32 32" "256 256"; do
  for v in 0 4 1 3 5 8 9; do
    timeout 60 ./umma_probe $v $shape 2>&1 | tail -8
  done
done | tee ../../gpurun_out/probe.log
