#!/bin/bash
# Round deliverables (1 GPU): parity tests, smoke, full bench lines (C3 headline, C2), ncu launch list, DRAM traffic of the
# row-update launches at full size, one ncu --set full capture (X and Y halves) on the 1/5-scale shape, top-N bench,
# wait profile.  Outputs -> gpurun_out/r02_*; the summaries under profiles/ are made from them by scripts/ncu_summary.py.
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))" >> gpurun_out/r02_pytest_gpu.log; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s)-S ))"; tail -3 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "bench c3 rc=$? t=$(( $(date +%s)-S ))"
cut -c1-1500 gpurun_out/r02_bench_c3.json; tail -3 gpurun_out/r02_bench_c3.err
timeout 600 python bench.py --config c2 --steps 10 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; echo "bench c2 rc=$? t=$(( $(date +%s)-S ))"; cut -c1-400 gpurun_out/r02_bench_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu launches rc=$? t=$(( $(date +%s)-S ))"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:row_update_v2 -s 6 -c 2 --csv --log-file gpurun_out/r02_dram_c3.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_dram.log 2>&1; echo "ncu dram rc=$? t=$(( $(date +%s)-S ))"; tail -4 gpurun_out/r02_dram_c3.csv | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_update_v2 -s 6 -c 2 -f -o gpurun_out/r02_prof_v2_c3p python bench.py --config c3p --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$? t=$(( $(date +%s)-S ))"
timeout 300 python scripts/topn_bench.py > gpurun_out/r02_topn.json 2> gpurun_out/r02_topn.err; echo "topn rc=$?"; cut -c1-600 gpurun_out/r02_topn.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:topn -c 60 --csv --log-file gpurun_out/r02_topn_launches.csv python scripts/topn_bench.py > /dev/null 2>&1; echo "ncu topn rc=$? t=$(( $(date +%s)-S ))"
timeout 300 python scripts/wait_profile.py c3p > gpurun_out/r02_wait_profile_c3p.txt 2>&1; tail -24 gpurun_out/r02_wait_profile_c3p.txt
timeout 300 python scripts/e2e_phases.py 2>&1 | tail -9 > gpurun_out/r02_e2e_phases.txt; cat gpurun_out/r02_e2e_phases.txt
echo "done t=$(( $(date +%s)-S ))"
