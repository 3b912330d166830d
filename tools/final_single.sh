#!/bin/bash
# one-GPU measurement set of the round: GPU tests, smoke, default bench (sustained + burst), the other configs
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_gputest.log 2>&1; tail -2 gpurun_out/r2_final_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 50 --warmup 3 > gpurun_out/r2_final_bench_1gpu.json 2>gpurun_out/r2_final_bench_1gpu.err; tail -c 400 gpurun_out/r2_final_bench_1gpu.json
python bench.py --steps 20 --warmup 3 --preheat 0 --no-cpu-baseline > gpurun_out/r2_final_bench_1gpu_burst.json 2>/dev/null
for c in 1 3 4 5; do python bench.py --config $c --steps 5 --warmup 3 --preheat 1 2>/dev/null; done > gpurun_out/r2_final_other_configs.jsonl
for b in 1 8 16 128 256 512; do python bench.py --config 3 --batch $b --stored 4000 --steps 4 --warmup 3 --preheat 0.5 2>/dev/null; done > gpurun_out/r2_final_cfg3_batches.jsonl
python - <<'PY'
import json
def show(f):
    for l in open(f):
        l=l.strip()
        if not l.startswith('{'): continue
        d=json.loads(l); r=d['roofline']
        print('%-60s B=%s ms %.3f value %.4g kern_frac %s whole %s e2e_ms %.3f' % (d['config']['workload'][:60], d['config'].get('batch','-'), d['ms_per_step'], d['value'], round(r['frac'],3) if r['frac'] else None, round(r.get('whole_step_frac') or 0,3), d['e2e']['ms_per_step']))
for f in ('gpurun_out/r2_final_bench_1gpu.json','gpurun_out/r2_final_bench_1gpu_burst.json','gpurun_out/r2_final_other_configs.jsonl','gpurun_out/r2_final_cfg3_batches.jsonl'): show(f)
PY
