#!/bin/bash
# N-GPU run: bench (both arms) + configs 4/5 sweep.  usage: tools/gpu/gpu_mgpu.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 \
    2> gpurun_out/bench_${N}gpu.err | tee gpurun_out/bench_${N}gpu.json | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tools/sweep_mgpu.py \
    --out gpurun_out/sweep_mgpu_$N.json > /dev/null 2> gpurun_out/sweep_mgpu_$N.err; echo "sweep rc=$?"
grep "^{'gates" gpurun_out/sweep_mgpu_$N.err | tail -5; grep -A4 adder gpurun_out/sweep_mgpu_$N.json | head -6
