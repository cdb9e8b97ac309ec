#!/bin/bash
# fetch look-ahead A/B on C3 and C2 (gated on parity)
mkdir -p gpurun_out
tag=s2f
for v in f2 f2k12 f2p11; do
  MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "warp_role or large_config or fixed_iterations or ragged or headline_cutdown" > gpurun_out/${tag}_gate_$v.log 2>&1; echo "gate $v rc=$? $(tail -1 gpurun_out/${tag}_gate_$v.log)"
done
echo "base:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/${tag}_c3_base.txt | cut -c1-170
for v in f1 f2; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/${tag}_c3_$v.txt | cut -c1-170; done
echo "base:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/${tag}_c3_base.txt | cut -c1-170
for rep in 1 2; do
  echo "base:"; ./scripts/quick_bench.sh c2 10 2>&1 | tee -a gpurun_out/${tag}_c2_base.txt | cut -c1-170
  for v in f1 f2 k12 f2k12 f2p11; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c2 10 2>&1 | tee -a gpurun_out/${tag}_c2_$v.txt | cut -c1-170; done
done
