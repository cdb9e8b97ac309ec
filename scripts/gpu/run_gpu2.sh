#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
cut -c1-1800 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_simt_c3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:row_update -s 6 -c 2 -o gpurun_out/prof_simt_c3 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
