#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 1 -c 1 -f -o gpurun_out/prof_br \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/plain.log | cut -c1-300
