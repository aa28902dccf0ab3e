#!/bin/bash
# GPU box: smoke, bench (N=1), then ncu launch list + one full capture of the dominant kernel and of the key switch
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 1 -c 1 -f -o gpurun_out/prof_br \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:keyswitch -s 1 -c 1 -f -o gpurun_out/prof_ks \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
echo "ncu3 rc=$?"
ls -la gpurun_out
