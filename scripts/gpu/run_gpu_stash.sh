#!/bin/bash
# stash mode (fp64 re-solve from the tensor-core data term): parity tests, c5p line, c3 regression check
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_parity_gpu.py -x -q -k "stash or split or ragged or powerlaw or warp_role or singular" > gpurun_out/s3d_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/s3d_pytest.log | cut -c1-300
timeout 120 python bench.py --config c5p --steps 5 --warmup 3 > gpurun_out/s3d_c5p.json 2> gpurun_out/s3d_c5p.err
echo "c5p rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/s3d_c5p.json').read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], r['avg_launch_ms'], r['other_half']['avg_launch_ms'], r['frac'], r['iteration_frac_of_hbm_roof'], r['fp64_retry_rows_per_iteration'], d['parity']['rank0'], d['parity']['ok'])"
BENCH_TIMEOUT=100 bash scripts/quick_bench.sh c3 3 2>&1 | tee gpurun_out/s3d_c3.txt
