#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python -c "
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'br_ms', round(d['kernels']['blind_rotate_ms'],3), 'ks_ms', round(d['kernels']['keyswitch_ms'],3), 'lat1_us', round(d['latency_us_single_gate']), 'wrong', d['wrong_bits'], 'clk', d['clocks'], 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('value_1core'))
"
python tools/sweep.py --out gpurun_out/sweeps.json > /dev/null 2> gpurun_out/sweep.err; echo "sweep rc=$?"; tail -3 gpurun_out/sweep.err
