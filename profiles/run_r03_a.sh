#!/bin/bash
mkdir -p gpurun_out/r03a
for prio in 1 0 1 0; do
SLNLP_STREAM_PRIO=$prio python bench.py --steps 100 --warmup 10 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('prio $prio: cfg1', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step'], 'e2e', round(d['e2e']['value']))"
done
SLNLP_STREAM_PRIO=1 python profiles/timeline_step.py cfg1 bf16 > gpurun_out/r03a/timeline_prio1.txt 2>&1
for wl in cfg2 cfg3; do for prio in 1 0; do
SLNLP_STREAM_PRIO=$prio python bench.py --workload $wl --steps 50 --warmup 10 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('prio $prio: $wl', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms')"
done; done
