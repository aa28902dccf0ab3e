#!/bin/bash
mkdir -p gpurun_out
for v in 0 2 3 5; do
  echo "== variant $v"
  TFHE_B200_BR_VARIANT=$v python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'br_ms', round(d['kernels']['blind_rotate_ms'],3), 'ks_ms', round(d['kernels']['keyswitch_ms'],3), 'lat', round(d['latency_us_single_gate']), 'wrong', d['wrong_bits'], 'clk', d['clocks']['sm_mhz'])
"
done
for v in 3 5; do TFHE_B200_BR_VARIANT=$v python -m pytest tests -m gpu -x -q 2>&1 | tail -2; done
