#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/s2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2j_pytest.log; tail -4 gpurun_out/s2j_pytest.log
timeout 300 python scripts/e2e_phases.py 2>&1 | tail -9 | tee gpurun_out/s2j_e2e_phases.txt
