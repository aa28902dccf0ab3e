#!/bin/bash
# bench line (N=1) + reference arm + launch list + quick extprod parity
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "external_product or cmux or golden" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; cat gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r2_bench_ref.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
