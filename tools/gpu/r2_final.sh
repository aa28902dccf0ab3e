#!/bin/bash
# end-of-milestone run: full GPU suite, smoke, bench line, reference arm, launch list, ncu capture of the dominant kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2_bench_ref.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 300 python tools/brtime.py 1184 > gpurun_out/r2_f64_prof_plain.log 2>&1; tail -1 gpurun_out/r2_f64_prof_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_f64 -s 1 -c 1 -f -o gpurun_out/prof_f64 \
    python tools/brtime.py 1184 > gpurun_out/r2_f64_prof_ncu.log 2>&1
echo "ncu full rc=$?"
