#!/bin/bash
# same-box A/B: base (in-tree lib) vs variants, interleaved twice
mkdir -p gpurun_out
for rep in 1 2; do
  echo "base:";   ./scripts/quick_bench.sh ${CFG:-c3} 3 2>&1 | tee -a gpurun_out/ab2_base.txt
  for v in "$@"; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh ${CFG:-c3} 3 2>&1 | tee -a gpurun_out/ab2_$v.txt; done
done
