#!/bin/bash
# per-rank share of the strong split on ONE GPU: step time at n = 1024 / G samples, for several samples-per-launch settings
for n in 1024 512 256 128; do
  for ns in 0 32 64 128; do
    if [ $ns -gt $n ]; then continue; fi
    DBX_FUSED_NS=$ns python bench.py --samples $n --steps 40 --warmup 3 --preheat 1.0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('n=%4d ns=%3s  ms/step %.4f  kernel_ms %.4f launches/step %.1f  e2e %.4f host_enq %.3f  clk %s' % ($n, '$ns', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['kernel_launches_per_step'], d['e2e']['ms_per_step'], d['config']['host_enqueue_ms_per_step'], d['clocks']['sm_mhz']))"
  done
done
