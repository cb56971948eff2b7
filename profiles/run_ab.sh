#!/bin/bash
# parity suite, then bench lines with and without programmatic dependent launch
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for wl in cfg1 cfg2 cfg3; do for pdl in 1 0; do
SLNLP_PDL=$pdl python bench.py --workload $wl --precision bf16 --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$wl pdl=$pdl', round(d['value']),'seq/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value']))"
done; done
SLNLP_PDL=1 python bench.py --workload cfg1 --precision fp32 --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cfg1 fp32', round(d['value']),'seq/s', round(d['ms_per_step'],3),'ms')"
