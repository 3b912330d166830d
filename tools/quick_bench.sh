#!/bin/bash
# usage: tools/quick_bench.sh [bench args...]  -> one summary line of bench.py (no CPU leg)
python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('ms/step %.4f kernel_ms %.4f achieved %.1f frac %.3f e2e %.4f clk %s %s' % (d['ms_per_step'], r['kernel_ms_per_step'], r['achieved'], r['frac'], d['e2e']['ms_per_step'] or -1, d['clocks']['sm_mhz'], d['clocks']['reasons']))"
