#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
