#!/bin/bash
# one pytest selection: scripts/gpu/run_gpu_one.sh "<-k expression>"
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q -k "$1" > gpurun_out/pytest_one.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_one.log; tail -12 gpurun_out/pytest_one.log
