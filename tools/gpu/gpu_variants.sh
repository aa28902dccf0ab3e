#!/bin/bash
# the GPU parity suite under every selectable kernel shape / mode (environment switches read by tfhe_b200_ctx_create)
mkdir -p gpurun_out
: > gpurun_out/variants.log
run() { echo "== $*" | tee -a gpurun_out/variants.log; env "$@" python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee -a gpurun_out/variants.log; }
run TFHE_B200_DEAL_FIXED=1
run TFHE_B200_BR_VARIANT=3
run TFHE_B200_BR_VARIANT=9
run TFHE_B200_PAIR_MAX=20
run TFHE_B200_KS_VARIANT=1
run TFHE_B200_KEY_SLICES=2
