#!/bin/bash
# ncu capture of the FFT64 blind rotation with ONE warp per scheduler (4 gates per SM): what a lone warp stalls on
mkdir -p gpurun_out
export TFHE_B200_KEY_SLICES=1
timeout 300 python tools/brtime.py ${PROF_B:-592} > gpurun_out/r2_f64_prof_plain.log 2>&1; tail -2 gpurun_out/r2_f64_prof_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_f64 -s 1 -c 1 -f -o gpurun_out/prof_f64_${PROF_B:-592} \
    python tools/brtime.py ${PROF_B:-592} > gpurun_out/r2_f64_prof_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_f64_prof_ncu.log
