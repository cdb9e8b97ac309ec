#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/s2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2e_pytest.log; tail -4 gpurun_out/s2e_pytest.log
timeout 200 python scripts/wait_profile.py c2 > gpurun_out/s2e_wait_c2.txt 2>&1; cat gpurun_out/s2e_wait_c2.txt | tail -26
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/s2e_topn_launches.csv python scripts/topn_bench.py 200000 64 10 > gpurun_out/s2e_topn_ncu.log 2>&1; echo "ncu rc=$?"
