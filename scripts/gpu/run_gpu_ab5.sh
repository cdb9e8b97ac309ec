#!/bin/bash
# the in-tree library must pass the parity gate first; then interleaved A/B on C3 and C2
mkdir -p gpurun_out
tag=$1; shift
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/${tag}_gate_base.log 2>&1; echo "gate base rc=$? $(tail -1 gpurun_out/${tag}_gate_base.log)"
for rep in 1 2; do
  echo "base:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/${tag}_c3_base.txt | cut -c1-170
  for v in "$@"; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/${tag}_c3_$v.txt | cut -c1-170; done
done
echo "base:"; ./scripts/quick_bench.sh c2 10 2>&1 | tee -a gpurun_out/${tag}_c2_base.txt | cut -c1-170
for v in "$@"; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c2 10 2>&1 | tee -a gpurun_out/${tag}_c2_$v.txt | cut -c1-170; done
