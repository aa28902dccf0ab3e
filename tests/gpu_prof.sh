#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 1 -c 1 -f -o gpurun_out/prof_br \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
for v in 3 5; do TFHE_B200_BR_VARIANT=$v python tools/sweep.py --max-log2 15 --out gpurun_out/sweeps_v$v.json > /dev/null 2>&1; python -c "
import json; d=json.load(open('gpurun_out/sweeps_v$v.json')); print('variant $v', [(r['gates'], round(r['gates_per_s'])) for r in d['config5_sweep_1gpu']], d['config4_adder32']['wall_seconds_1gpu'])"; done
