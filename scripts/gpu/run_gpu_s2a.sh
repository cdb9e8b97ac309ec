#!/bin/bash
# session-2 checkpoint: parity tests, quick C3 / C2 lines, one ncu --set full capture (X and Y halves, 1/5-scale shape)
mkdir -p gpurun_out
S=$(date +%s)
timeout 500 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/s2a_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))" >> gpurun_out/s2a_pytest.log; tail -6 gpurun_out/s2a_pytest.log
./scripts/quick_bench.sh c3 3 2>&1 | tee gpurun_out/s2a_c3.txt | cut -c1-260
./scripts/quick_bench.sh c2 5 2>&1 | tee gpurun_out/s2a_c2.txt | cut -c1-260
MYRRIX_ALS_V1=1 ./scripts/quick_bench.sh c2 5 2>&1 | tee gpurun_out/s2a_c2_v1.txt | cut -c1-260
timeout 400 ncu --set full --clock-control none --import-source on -k regex:row_update_v2 -s 6 -c 2 -f -o gpurun_out/s2a_prof_v2_c3p python bench.py --config c3p --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/s2a_ncu_full.log 2>&1; echo "ncu full rc=$? t=$(( $(date +%s)-S ))"
