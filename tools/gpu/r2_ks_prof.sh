#!/bin/bash
# ncu capture of the key-switch kernel at 1024 gates
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:keyswitch_p -s 1 -c 1 -f -o gpurun_out/prof_ks \
    python tools/brtime.py 1024 > gpurun_out/r2_ks_prof_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_ks_prof_ncu.log
