#!/bin/bash
# GPU parity suite + gemm microbench + bench lines of the tensor-core path for cfg1/cfg2/cfg3
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for wl in cfg1 cfg2 cfg3; do
python bench.py --workload $wl --precision bf16 --steps 30 --warmup 3 --no-cpu-baseline 2>gpurun_out/b_${wl}_bf16.err > gpurun_out/b_${wl}_bf16.json; echo "$wl rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/b_${wl}_bf16.json').read().strip().splitlines()[-1]);print('$wl', round(d['value']),'seq/s', round(d['ms_per_step'],3),'ms', d['launches_per_step'],'launches e2e', round(d['e2e']['value']))"
done
