#!/bin/bash
# FFT64 kernel: start-up stagger sweep (and any experiment libraries under rustfhe_b200/exp/)
mkdir -p gpurun_out
export TFHE_B200_KEY_SLICES=1
: > gpurun_out/r2_f64_sweep.log
for sg in ${STAGGERS:-0 200 400 800 1600}; do
  echo "== stagger $sg" | tee -a gpurun_out/r2_f64_sweep.log
  TFHE_B200_F64_STAGGER=$sg timeout 300 python tools/brtime.py ${F64_SIZES:-1184 7104} 2>&1 | tail -2 | tee -a gpurun_out/r2_f64_sweep.log
done
for lib in rustfhe_b200/exp/lib_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib" | tee -a gpurun_out/r2_f64_sweep.log
  TFHE_B200_LIB=$PWD/$lib timeout 300 python tools/brtime.py ${F64_SIZES:-1184 7104} 2>&1 | tail -2 | tee -a gpurun_out/r2_f64_sweep.log
done
