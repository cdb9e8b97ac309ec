#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_topn.py -m gpu -q -x --timeout 200 > gpurun_out/s2c_pytest_topn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2c_pytest_topn.log; tail -25 gpurun_out/s2c_pytest_topn.log
timeout 300 python scripts/topn_bench.py > gpurun_out/s2c_topn.json 2> gpurun_out/s2c_topn.err; echo "topn rc=$?"; cat gpurun_out/s2c_topn.json; tail -3 gpurun_out/s2c_topn.err
