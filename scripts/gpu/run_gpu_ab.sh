#!/bin/bash
# same-box A/B of the in-tree library ("base") against scripts/_var/<variant>.so, then the wait profile of scripts/_prof
# usage: run_gpu_ab.sh TAG variant...
mkdir -p gpurun_out
tag=$1; shift
echo "base:"; ./scripts/quick_bench.sh ${CFG:-c3} 3 2>&1 | tee -a gpurun_out/${tag}_ab_base.txt
for v in "$@"; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh ${CFG:-c3} 3 2>&1 | tee -a gpurun_out/${tag}_ab_$v.txt; done
echo "base:"; ./scripts/quick_bench.sh ${CFG:-c3} 3 2>&1 | tee -a gpurun_out/${tag}_ab_base.txt
if [ -z "$NOPROF" ]; then echo "== wait profile"; timeout 300 python scripts/wait_profile.py c3p > gpurun_out/${tag}_wait_profile.txt 2>&1; tail -26 gpurun_out/${tag}_wait_profile.txt; fi
