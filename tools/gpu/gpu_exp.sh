#!/bin/bash
# experiment runner: every rustfhe_b200/exp/lib_*.so through tools/brtime.py (timing experiments; wrong bits expected for EXP_ builds)
mkdir -p gpurun_out
: > gpurun_out/exp.log
for lib in ${EXP_LIBS:-rustfhe_b200/exp/lib_*.so}; do
  for v in ${EXP_VARIANTS:-3}; do
    TFHE_B200_LIB=$PWD/$lib TFHE_B200_BR_VARIANT=$v timeout 300 python tools/brtime.py ${EXP_SIZES:-1024 7104} 2>&1 | tail -4 | tee -a gpurun_out/exp.log
  done
done
