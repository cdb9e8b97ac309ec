#!/bin/bash
# Round-1 deliverables for the current tcgen05 kernel: parity tests, full bench line, ncu launch list, one ncu --set full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
S=$(date +%s)
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$? t=$(( $(date +%s)-S ))"
cut -c1-2500 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tcgen05_c3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$? t=$(( $(date +%s)-S ))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_update_umma -s 6 -c 1 -o gpurun_out/prof_umma_v5_c3p python bench.py --config c3p --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$? t=$(( $(date +%s)-S ))"
timeout 300 python scripts/wait_profile.py c3p > gpurun_out/wait_profile.log 2>&1; tail -25 gpurun_out/wait_profile.log
ls -la gpurun_out
