#!/bin/bash
# build an experiment variant of the library:  tools/build_exp.sh NAME [-DFLAG ...]   -> rustfhe_b200/exp/lib_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p rustfhe_b200/exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-pthread -shared -cudart static "$@" \
    rustfhe_b200/csrc/engine.cu rustfhe_b200/csrc/group.cu rustfhe_b200/csrc/hostkeys.cpp rustfhe_b200/csrc/wire.cpp -lnccl -o rustfhe_b200/exp/lib_$name.so
echo built rustfhe_b200/exp/lib_$name.so
