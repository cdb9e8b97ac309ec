#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_topn.py -m gpu -q -x --timeout 200 2>&1 | tail -3
timeout 300 python scripts/topn_bench.py > gpurun_out/s2l_topn.json 2> gpurun_out/s2l_topn.err; echo "topn rc=$?"; cut -c1-700 gpurun_out/s2l_topn.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:topn -c 40 --csv --log-file gpurun_out/s2l_topn_launches.csv python scripts/topn_bench.py > /dev/null 2>&1; grep -E "topn" gpurun_out/s2l_topn_launches.csv | tail -4 | cut -d'"' -f10,28,30
