#!/bin/bash
# end of round 2: the whole GPU suite, smoke, the c5p line
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02b_pytest_gpu.log; tail -3 gpurun_out/r02b_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 100 python bench.py --config c5p --steps 5 --warmup 3 > gpurun_out/r02b_bench_c5p.json 2> gpurun_out/r02b_bench_c5p.err
echo "c5p rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02b_bench_c5p.json').read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], r['avg_launch_ms'], r['other_half']['avg_launch_ms'], r['frac'], r['iteration_frac_of_hbm_roof'], d['parity']['ok'], d['parity']['worst_over_ranks'])"
