#!/bin/bash
# long-row split: parity tests, then c5p with the split on (default limit) and off
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_parity_gpu.py -x -q -k "split or ragged or powerlaw or warp_role" > gpurun_out/s3a_pytest_split.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/s3a_pytest_split.log
timeout 100 python bench.py --config c5p --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s3a_c5p_split.json 2> gpurun_out/s3a_c5p_split.err
echo "c5p split rc=$?"; tail -c 1500 gpurun_out/s3a_c5p_split.json
MYRRIX_ALS_SPLIT_ROWS=0 BENCH_TIMEOUT=70 bash scripts/quick_bench.sh c5p 3 > gpurun_out/s3a_c5p_nosplit.txt 2>&1
cat gpurun_out/s3a_c5p_nosplit.txt
