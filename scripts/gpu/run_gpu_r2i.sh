#!/bin/bash
mkdir -p gpurun_out
echo "== v2_debug (phase profile build)"; MYRRIX_ALS_LIB=$PWD/scripts/_var/sprof.so timeout 240 python scripts/v2_debug.py 2>&1 | head -4
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -4 gpurun_out/r2i_pytest.log
echo "== c10 correctness"; MYRRIX_ALS_LIB=$PWD/scripts/_var/c10.so timeout 600 python -m pytest tests -m gpu -q --timeout 400 -k "headline or warp_role or ragged or fixed_iterations" 2>&1 | tail -4
NOPROF=1 bash scripts/gpu/run_gpu_ab.sh r2i "$@"
