#!/bin/bash
# gpurun with retries on "busy / no slot" (exit 3): tools/gpu/run.sh TIMEOUT 'command' [log]
t=$1; cmd=$2; log=${3:-/tmp/gpurun.log}
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$cmd" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "rc=$rc" >> "$log"; exit $rc; fi
  sleep 60
done
echo "gave up" >> "$log"; exit 3
