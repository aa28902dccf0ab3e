#!/bin/bash
# gpurun --gpus N with retries: tools/gpu/run_n.sh N TIMEOUT 'command' [log]
n=$1; t=$2; cmd=$3; log=${4:-/tmp/gpurun_n.log}
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus "$n" --timeout "$t" -- "$cmd" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "rc=$rc" >> "$log"; exit $rc; fi
  sleep 90
done
echo "gave up" >> "$log"; exit 3
