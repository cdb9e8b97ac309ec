#!/bin/bash
# Gate, then A/B.  usage: run_gpu_gate_ab.sh TAG "gate variants (watchdog builds)" variant...
# A variant is benchmarked only after every gate variant passed the k = 64 parity tests within 150 s.
mkdir -p gpurun_out
tag=$1; gates=$2; shift; shift
ok=1
for v in $gates; do
  echo "== gate $v"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so timeout 150 python -m pytest tests -m gpu -q -x -s -k "fixed_iterations and 1200 or warp_role or headline_cutdown or blocked_solver" > gpurun_out/${tag}_gate_$v.log 2>&1; rc=$?
  grep -E "WATCHDOG" gpurun_out/${tag}_gate_$v.log | sort | uniq -c | head -12; tail -2 gpurun_out/${tag}_gate_$v.log; echo "rc=$rc"
  if [ $rc -ne 0 ]; then ok=0; fi
done
if [ $ok -eq 1 ]; then NOPROF=1 bash scripts/gpu/run_gpu_ab.sh $tag "$@"; else echo "gate failed: no benchmark"; fi
