#!/bin/bash
# N-GPU regression after a change to the step: independent fits, data-parallel all-reduce (cfg1 model,
# global batch 50 N), the cfg5 grid slice farmed over N GPUs.
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1], round(d['value'],1), d['unit'], 'n_gpus', d['n_gpus'], 'ms', round(d.get('ms_per_step') or 0,3))" $1; }
$TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/multi_fits_n$N.json 2> gpurun_out/multi_fits_n$N.err; echo "fits rc=$?"; show gpurun_out/multi_fits_n$N.json
$TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --dp --batch $((50 * N)) > gpurun_out/multi_dp_cfg1_n$N.json 2> gpurun_out/multi_dp_cfg1_n$N.err; echo "dp cfg1 rc=$?"; show gpurun_out/multi_dp_cfg1_n$N.json; tail -2 gpurun_out/multi_dp_cfg1_n$N.err
$TR bench.py --gpus $N --workload cfg5 --grid-fraction 0.2 > gpurun_out/multi_grid_n$N.json 2> gpurun_out/multi_grid_n$N.err; echo "grid rc=$?"; show gpurun_out/multi_grid_n$N.json
