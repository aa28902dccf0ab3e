#!/bin/bash
mkdir -p gpurun_out
export TFHE_B200_F64_LATENCY=2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:latency2 -s 1 -c 1 -f -o gpurun_out/prof_l2 \
    python tools/brtime.py 8 > gpurun_out/r2_l2_prof_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_l2_prof_ncu.log
