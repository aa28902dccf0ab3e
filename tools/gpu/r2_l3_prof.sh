#!/bin/bash
# ncu capture of the 2-SM cluster latency kernel at 8 gates
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:latency3 -s 1 -c 1 -f -o gpurun_out/prof_l3 \
    python tools/brtime.py 8 > gpurun_out/r2_l3_prof_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_l3_prof_ncu.log
