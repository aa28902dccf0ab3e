#!/bin/bash
# quick timing of the default build (+ any experiment libs) and a subset of the parity tests
mkdir -p gpurun_out
timeout 300 python tools/brtime.py 888 1024 7104 2>&1 | tail -3 | tee gpurun_out/r2_quick.log
for lib in rustfhe_b200/exp/lib_*.so; do
  [ -f "$lib" ] || continue
  TFHE_B200_LIB=$PWD/$lib timeout 300 python tools/brtime.py 888 7104 2>&1 | tail -2 | tee -a gpurun_out/r2_quick.log
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${QUICK_TESTS:-truth or launch_shapes or batch_1024 or exact_mode or blind_rotate}" 2>&1 | tail -5 | tee -a gpurun_out/r2_quick.log
