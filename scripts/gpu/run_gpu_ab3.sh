#!/bin/bash
# gate (parity) then interleaved A/B on C3 and C2: run_gpu_ab3.sh TAG variant...
mkdir -p gpurun_out
tag=$1; shift
for v in "$@"; do
  MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "warp_role or large_config or fixed_iterations or ragged or headline_cutdown or variants" > gpurun_out/${tag}_gate_$v.log 2>&1; echo "gate $v rc=$? $(tail -1 gpurun_out/${tag}_gate_$v.log)"
done
for rep in 1 2; do
  echo "base:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/${tag}_c3_base.txt | cut -c1-170
  for v in "$@"; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/${tag}_c3_$v.txt | cut -c1-170; done
done
echo "base:"; ./scripts/quick_bench.sh c2 10 2>&1 | tee -a gpurun_out/${tag}_c2_base.txt | cut -c1-170
for v in "$@"; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c2 10 2>&1 | tee -a gpurun_out/${tag}_c2_$v.txt | cut -c1-170; done
