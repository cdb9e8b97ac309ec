#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/s2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2n_pytest.log; tail -5 gpurun_out/s2n_pytest.log
for rep in 1 2; do
echo "c5p ordered:"; timeout 300 python bench.py --config c5p --steps 3 --warmup 2 2>gpurun_out/s2n_c5p.err | tee gpurun_out/s2n_c5p_order.json | python scripts/e2e_pick.py 2>/dev/null; ./scripts/quick_bench.sh c5p 3 | cut -c1-200
echo "c5p as stored:"; MYRRIX_ALS_NO_ROW_ORDER=1 ./scripts/quick_bench.sh c5p 3 | cut -c1-200
done
echo "c3:"; ./scripts/quick_bench.sh c3 3 | cut -c1-200
echo "c3 as stored:"; MYRRIX_ALS_NO_ROW_ORDER=1 ./scripts/quick_bench.sh c3 3 | cut -c1-200
