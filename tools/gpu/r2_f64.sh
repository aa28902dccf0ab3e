#!/bin/bash
# FFT64 mode (TFHE_B200_KEY_SLICES=1): timing of the one-warp-per-gate kernel, then the parity tests that exercise large batches
mkdir -p gpurun_out
export TFHE_B200_KEY_SLICES=1
timeout 300 python tools/brtime.py ${F64_SIZES:-1184 1024 2368 7104} 2>&1 | tail -5 | tee gpurun_out/r2_f64.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${QUICK_TESTS:-truth or launch_shapes or batch_1024 or blind_rotate or golden or mixed}" 2>&1 | tail -8 | tee -a gpurun_out/r2_f64.log
