#!/bin/bash
# data-parallel step at N GPUs: captured (default) vs eager (SLNLP_DP_GRAPH=0); cfg1 model, global batch 50 N
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2], round(d['value'],1), d['unit'], 'n_gpus', d['n_gpus'], 'ms', round(d.get('ms_per_step') or 0,3), 'e2e', round(d['e2e']['value']))" $1 "$2"; }
for g in 1 0; do
SLNLP_DP_GRAPH=$g timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --dp --batch $((50 * N)) > gpurun_out/multi_dp_cfg1_n${N}_g$g.json 2> gpurun_out/multi_dp_cfg1_n${N}_g$g.err; echo "dp cfg1 graph=$g rc=$?"; show gpurun_out/multi_dp_cfg1_n${N}_g$g.json "graph=$g" || tail -5 gpurun_out/multi_dp_cfg1_n${N}_g$g.err
done
