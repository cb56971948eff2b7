#!/bin/bash
# round 2, eight GPUs: the driver's N = 8 command (replicas + dp + grid sub-records), then N = 4
mkdir -p gpurun_out/r02n8
nvidia-smi -L | wc -l
for n in 8 4; do
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 ) > gpurun_out/r02n8/bench_n$n.json 2> gpurun_out/r02n8/bench_n$n.err; echo "bench n$n rc=$?"
grep -v "Warn\|warn\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r02n8/bench_n$n.err | tail -6
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r02n8/bench_n$n.json').read().strip().splitlines()[-1])
    print('N=$n replicas', round(d['value']), 'seq/s; dp', round(d['dp']['value']), 'seq/s', round(d['dp']['ms_per_step'],2), 'ms', d['dp'].get('allreduce',{}).get('busbw_gbs'), d['dp'].get('dp_parity'), '; grid', round(d['grid']['value']), 'fits/h', round(d['grid']['search_seconds'],1), 's')
except Exception as e: print('ERR', e)
P
done
