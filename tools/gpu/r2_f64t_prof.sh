#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_f64t -s 1 -c 1 -f -o gpurun_out/prof_f64t \
    python tools/brtime.py 1776 > gpurun_out/r2_f64t_prof_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_f64t_prof_ncu.log
