#!/bin/bash
mkdir -p gpurun_out/r03n4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 --legs dp,grid > gpurun_out/r03n4/bench_n4.json 2> gpurun_out/r03n4/bench_n4.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03n4/bench_n4.json').read().strip().splitlines()[-1])
print('n8 headline', round(d['value']), d['unit'], d['ms_per_step'])
dp=d.get('dp') or {}
print('dp', {k: dp.get(k) for k in ('value','ms_per_step','exposed_exchange_ms','limiter','dp_parity','error')})
print('allreduce', dp.get('allreduce'))
g=d.get('grid') or {}
print('grid', {k: g.get(k) for k in ('value','search_seconds','fits','error')})
PY
