#!/bin/bash
mkdir -p gpurun_out/r03f
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03f/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r03f/smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --legs dp > gpurun_out/r03f/bench_n2.json 2> gpurun_out/r03f/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03f/bench_n2.json').read().strip().splitlines()[-1])
print('n2 headline', round(d['value']), d['unit'], d['ms_per_step'])
dp=d.get('dp') or {}
print('dp', {k: dp.get(k) for k in ('value','ms_per_step','exposed_exchange_ms','limiter','dp_parity','error')})
print('allreduce', dp.get('allreduce'))
PY
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout=500 2>&1 | tail -2
