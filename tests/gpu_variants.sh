#!/bin/bash
mkdir -p gpurun_out
for v in 0 2 3; do
  echo "== variant $v"
  TFHE_B200_BR_VARIANT=$v python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'br_ms', d['kernels']['blind_rotate_ms'], 'ks_ms', d['kernels']['keyswitch_ms'], 'wrong', d['wrong_bits'], 'clk', d['clocks'])
"
done
TFHE_B200_BR_VARIANT=3 python -m pytest tests -m gpu -x -q -k "blind_rotate_exact or batch_1024 or truth" 2>&1 | tail -3
TFHE_B200_BR_VARIANT=2 python -m pytest tests -m gpu -x -q -k "blind_rotate_exact or batch_1024" 2>&1 | tail -3
