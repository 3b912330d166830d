#!/bin/bash
# the driver's scaling series on one 8-GPU box: bench.py at N = 1, 2, 4, 8 (reference arm first at N = 1), lines -> gpurun_out/r2_scale_N.json
STEPS=${STEPS:-20}
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_ref_1.json 2>/dev/null
python bench.py --gpus 1 --steps $STEPS --warmup 3 > gpurun_out/r2_scale_1.json 2>gpurun_out/r2_scale_1.err
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps $STEPS --warmup 3 > gpurun_out/r2_scale_$n.json 2>gpurun_out/r2_scale_$n.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29699 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_ref_8.json 2>/dev/null
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open('gpurun_out/r2_scale_%d.json'%n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n,'FAILED',e); continue
    if n==1: base=d['value']
    print('N=%d value %.4g ms %.4f speedup %.3f eff %.3f e2e_ms %.4f frac %.3f coll %s shard %s clk %s' % (n, d['value'], d['ms_per_step'], d['value']/base, d['value']/base/n, d['e2e']['ms_per_step'], d['roofline']['frac'], d['config']['collective'][:22], d.get('shard_check'), d['clocks'].get('sm_mhz')))
for n in (1,8):
    try:
        d=json.loads(open('gpurun_out/r2_ref_%d.json'%n).read().strip().splitlines()[-1]); print('ref N=%d value %.4g cores %s'%(n,d['value'],d['cpu_baseline']['cores']))
    except Exception as e: print('ref',n,'FAILED',e)
PY
