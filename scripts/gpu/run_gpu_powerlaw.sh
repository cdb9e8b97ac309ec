#!/bin/bash
# Power-law path (1 GPU): parity tests of the chunked rows / stash mode / sharded generator, the c5p bench line,
# same-box A/B of the switches (split limit, stash mode), ncu launch list of the row updates.
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_parity_gpu.py -x -q -k "stash or split or ragged or powerlaw or warp_role or singular" > gpurun_out/pl_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pl_pytest.log
timeout 120 python bench.py --config c5p --steps 5 --warmup 3 > gpurun_out/pl_bench_c5p.json 2> gpurun_out/pl_bench_c5p.err; echo "c5p rc=$?"
MYRRIX_ALS_STASH=0 BENCH_TIMEOUT=60 bash scripts/quick_bench.sh c5p 5 2>&1 | tee gpurun_out/pl_c5p_nostash.txt
for lim in 0 2048 16384; do
  echo "split limit $lim"; MYRRIX_ALS_SPLIT_ROWS=$lim BENCH_TIMEOUT=60 bash scripts/quick_bench.sh c5p 5 2>&1 | tee gpurun_out/pl_c5p_lim$lim.txt
done
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"row_update|resolve" -c 18 --csv --log-file gpurun_out/pl_launches_c5p.csv python bench.py --config c5p --steps 1 --warmup 3 --no-parity > gpurun_out/pl_ncu.log 2>&1; echo "ncu rc=$?"
