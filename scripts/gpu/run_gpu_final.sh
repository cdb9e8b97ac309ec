#!/bin/bash
# Round deliverables: parity tests, smoke, full bench line, ncu launch list, DRAM traffic of the row-update launches
# at full C3 size, one ncu --set full capture (X and Y halves) on the 1/5-scale shape.
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s)-S ))"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$? t=$(( $(date +%s)-S ))"
cut -c1-1200 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$? t=$(( $(date +%s)-S ))"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:row_update_umma -s 6 -c 2 --csv --log-file gpurun_out/dram_c3.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_dram.log 2>&1; echo "ncu dram rc=$? t=$(( $(date +%s)-S ))"; tail -4 gpurun_out/dram_c3.csv | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_update_umma -s 6 -c 2 -f -o gpurun_out/prof_umma_final_c3p python bench.py --config c3p --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$? t=$(( $(date +%s)-S ))"
./scripts/quick_bench.sh c2 5 2>&1 | cut -c1-160
