#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_foldin.py -m gpu -q -x --timeout 200 > gpurun_out/s2i_pytest_foldin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2i_pytest_foldin.log; tail -25 gpurun_out/s2i_pytest_foldin.log
