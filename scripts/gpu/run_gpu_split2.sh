#!/bin/bash
# sharded power-law generator test, c5p line with popular rows in the parity sample, split-limit sweep
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_parity_gpu.py -x -q -k "sharded_powerlaw or sharded_synthesis or powerlaw_generator" > gpurun_out/s3b_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/s3b_pytest.log
timeout 120 python bench.py --config c5p --steps 3 --warmup 3 > gpurun_out/s3b_c5p.json 2> gpurun_out/s3b_c5p.err
echo "c5p rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/s3b_c5p.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['avg_launch_ms'], d['roofline']['other_half']['avg_launch_ms'], d['parity'])"
for lim in 1024 2048 4096 16384; do
  echo "limit $lim"; MYRRIX_ALS_SPLIT_ROWS=$lim BENCH_TIMEOUT=60 bash scripts/quick_bench.sh c5p 3 2>&1 | tee gpurun_out/s3b_c5p_lim$lim.txt
done
