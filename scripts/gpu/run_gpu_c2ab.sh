#!/bin/bash
# C2 (k = 32) A/B of role mixes: gate each variant on the k = 32 parity tests, then quick C2 lines
mkdir -p gpurun_out
tag=$1; shift
for v in "$@"; do
  MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so timeout 150 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "warp_role or large_config or (fixed_iterations and 32)" > gpurun_out/${tag}_gate_$v.log 2>&1; echo "gate $v rc=$? $(tail -1 gpurun_out/${tag}_gate_$v.log)"
done
for rep in 1 2; do
  echo "base:"; ./scripts/quick_bench.sh c2 10 2>&1 | tee -a gpurun_out/${tag}_c2_base.txt | cut -c1-170
  for v in "$@"; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c2 10 2>&1 | tee -a gpurun_out/${tag}_c2_$v.txt | cut -c1-170; done
done
