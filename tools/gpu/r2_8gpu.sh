#!/bin/bash
# 8-GPU run with the final code: group tests, bench line under torchrun, config 4 on 1 and 8 GPUs
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_group.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_group_${N}gpu_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 \
    2> gpurun_out/bench_${N}gpu.err > gpurun_out/bench_${N}gpu.json; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench_${N}gpu.json
timeout 600 python tools/gpu/group_circuit_bench.py 16 64 256 2>&1 | grep -v "^NCCL" | tee gpurun_out/r2_group_circuit_${N}gpu.jsonl
