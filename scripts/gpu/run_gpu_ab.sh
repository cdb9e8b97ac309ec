#!/bin/bash
# A/B variants + wait profile + ncu full capture (X and Y halves) on the 1/5-scale headline shape
mkdir -p gpurun_out
S=$(date +%s)
echo "base:";   ./scripts/quick_bench.sh c3p 3 2>&1 | tee gpurun_out/ab_base.txt
for v in "$@"; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c3p 3 2>&1 | tee gpurun_out/ab_$v.txt; done
echo "t=$(( $(date +%s)-S ))"
timeout 300 python scripts/wait_profile.py c3p > gpurun_out/wait_profile.log 2>&1; cat gpurun_out/wait_profile.log
echo "t=$(( $(date +%s)-S ))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_update_umma -s 6 -c 2 -f -o gpurun_out/prof_umma_cur_c3p python bench.py --config c3p --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$? t=$(( $(date +%s)-S ))"
