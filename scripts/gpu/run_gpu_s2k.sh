#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2 3; do
  echo "base:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/s2k_c3_base.txt | cut -c1-170
  echo "nopool:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/nopool.so ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/s2k_c3_nopool.txt | cut -c1-170
done
