#!/bin/bash
# round 2, first contact of the two-warps-per-gate throughput kernel: parity suite, then timing of the launch variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/r2_first.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2_pytest_gpu.log
for cfg in "6 1" "6 0" "6 3" "4 1" "4 3" "4 0"; do
  set -- $cfg
  echo "== T2 G=$1 TWREG=$2" | tee -a gpurun_out/r2_first.log
  TFHE_B200_T2_G=$1 TFHE_B200_T2_TWREG=$2 timeout 300 python tools/brtime.py 888 1024 1776 7104 2>&1 | tail -5 | tee -a gpurun_out/r2_first.log
done
echo "== old 6-warp kernel, two slices (variant 8)" | tee -a gpurun_out/r2_first.log
TFHE_B200_BR_VARIANT=8 timeout 300 python tools/brtime.py 1024 7104 2>&1 | tail -3 | tee -a gpurun_out/r2_first.log
echo "== three slices" | tee -a gpurun_out/r2_first.log
TFHE_B200_KEY_SLICES=3 timeout 300 python tools/brtime.py 1024 7104 2>&1 | tail -3 | tee -a gpurun_out/r2_first.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/r2_bench_first.json
