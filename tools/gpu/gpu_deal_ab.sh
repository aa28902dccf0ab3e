#!/bin/bash
# A/B of the two ways a full batch is cut into CTAs (dealt evenly / 4-gate CTAs only / decided per call) under the bench's stream overlap
mkdir -p gpurun_out
: > gpurun_out/deal_ab.log
for fx in -1 0 1; do for ns in 1 2; do
  echo "deal_fixed=$fx streams=$ns" >> gpurun_out/deal_ab.log
  TFHE_B200_DEAL_FIXED=$fx BENCH_STREAMS=$ns python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d.get('value_serial'))" >> gpurun_out/deal_ab.log
done; done
cat gpurun_out/deal_ab.log
