#!/bin/bash
# ncu capture of the throughput blind rotation (one full wave: 6 gates per SM) + experiment builds
mkdir -p gpurun_out
timeout 300 python tools/brtime.py 888 > gpurun_out/r2_prof_plain.log 2>&1; tail -2 gpurun_out/r2_prof_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_t2 -s 1 -c 1 -f -o gpurun_out/prof_t2 \
    python tools/brtime.py 888 > gpurun_out/r2_prof_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_prof_ncu.log
for lib in rustfhe_b200/exp/lib_*.so; do
  [ -f "$lib" ] || continue
  TFHE_B200_LIB=$PWD/$lib timeout 300 python tools/brtime.py 888 7104 2>&1 | tail -2 | tee -a gpurun_out/r2_exp.log
done
ls -la gpurun_out/prof_t2.ncu-rep
