#!/bin/bash
mkdir -p gpurun_out
export TFHE_B200_F64_TMEM=2 TFHE_B200_F64_LATENCY=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_f64w2 -s 1 -c 1 -f -o gpurun_out/prof_f64w2 \
    python tools/brtime.py 888 > gpurun_out/r2_f64w2_prof_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_f64w2_prof_ncu.log
