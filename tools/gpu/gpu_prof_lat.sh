#!/bin/bash
mkdir -p gpurun_out
python tools/brtime.py 8 > gpurun_out/lat_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 2 -c 1 -f -o gpurun_out/prof_lat python tools/brtime.py 8 > gpurun_out/ncu_lat.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/lat_plain.log
