#!/bin/bash
# compute-sanitizer logs of every launch shape (kept under profiles/ by the builder)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  SANITIZE_FULL_GATE=$([ $tool = memcheck ] && echo 1 || echo 0) timeout 1500 compute-sanitizer --tool $tool --print-limit 20 \
      python tools/gpu/sanitize.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/r02_sanitizer_$tool.log
done
