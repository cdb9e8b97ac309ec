#!/bin/bash
# quick device-resident bench line (no e2e / cpu baseline), prints the key numbers
cfg=${1:-c3}; steps=${2:-3}
timeout ${BENCH_TIMEOUT:-120} python bench.py --config $cfg --steps $steps --warmup 3 --no-cpu-baseline --no-e2e --no-parity 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%s kernel=%s it/s=%.3f ms/step=%.1f dom=%s %.1f ms other=%.1f ms gram=%.2f ms frac=%.3f iterfrac=%.3f retry=%s launches=%d clocks=%s' % (
 d['config']['workload'][:12], d['config']['kernel'], d['value'], d['ms_per_step'], r['kernel'][-10:], r['avg_launch_ms'], r['other_half']['avg_launch_ms'],
 r['gramian_ms_per_iteration'], r['frac'], r['iteration_frac_of_hbm_roof'], r.get('fp64_retry_rows_per_iteration'), d['gpu_launches'], d['clocks']))"
