#!/bin/bash
# round 2, call B: where does the second-generation X-half spend its time? A/B of slot / lockstep variants,
# wait profile, ncu full capture (1/5-scale shape) for the per-role breakdown
mkdir -p gpurun_out
echo "base:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/r2b_ab_base.txt
for v in nolock x2 x2nolock; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/r2b_ab_$v.txt; done
echo "base:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/r2b_ab_base.txt
echo "== wait profile"; timeout 300 python scripts/wait_profile.py c3p > gpurun_out/r2b_wait_profile.txt 2>&1; cat gpurun_out/r2b_wait_profile.txt | tail -30
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:row_update_v2 -s 6 -c 2 -f -o gpurun_out/r2b_prof_v2_c3p python bench.py --config c3p --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2b_ncu_full.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2b_ncu_full.log
