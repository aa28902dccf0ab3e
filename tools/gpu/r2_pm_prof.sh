#!/bin/bash
# ncu capture of the FFT64 negacyclic product kernel at batch 65536 (the launches of tools/sweep_config3.py: 10 per batch size)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:polymul_f64 -s 45 -c 1 -f -o gpurun_out/prof_pm \
    python tools/sweep_config3.py > gpurun_out/r2_pm_prof_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_pm_prof_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:external_product_f64 -s 35 -c 1 -f -o gpurun_out/prof_xp \
    python tools/sweep_config3.py > gpurun_out/r2_xp_prof_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_xp_prof_ncu.log
