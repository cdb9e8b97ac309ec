#!/bin/bash
# compute-sanitizer on small cases: memcheck on everything, racecheck (shared-memory hazards) per kernel family
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitizer_case.py all > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$? t=$(( $(date +%s)-S ))"; tail -4 gpurun_out/r02_sanitizer_memcheck.log
for k in 64 32 16; do
  timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python scripts/sanitizer_case.py $k > gpurun_out/r02_sanitizer_racecheck_k$k.log 2>&1; echo "racecheck k=$k rc=$? t=$(( $(date +%s)-S ))"; tail -4 gpurun_out/r02_sanitizer_racecheck_k$k.log
done
