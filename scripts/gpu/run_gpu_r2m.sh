#!/bin/bash
mkdir -p gpurun_out
for v in c10g c12g; do
echo "== $v correctness"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so timeout 200 python -m pytest tests -m gpu -q -x --timeout 100 -k "fixed_iterations or warp_role or ragged or headline_config" 2>&1 | tail -2
done
NOPROF=1 bash scripts/gpu/run_gpu_ab.sh r2m "$@"
