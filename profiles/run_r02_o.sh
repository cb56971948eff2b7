#!/bin/bash
mkdir -p gpurun_out/r02o
python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/r02o/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02o/pytest.log
for sk in 0 1; do
SLNLP_SKINNY=$sk python bench.py --steps 100 --warmup 10 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('skinny $sk: cfg1', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step'], 'e2e', round(d['e2e']['value']))"
done
SLNLP_RNN_FUSED_DROPOUT=0 python bench.py --steps 100 --warmup 10 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('no mask-dropout: cfg1', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step'])"
python bench.py --steps 50 --warmup 10 --precision fp32 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('fp32', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step'])"
for wl in cfg2 cfg3; do
python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$wl', round(d['value']), 'seq/s', round(d['ms_per_step'],4), 'launches', d['launches_per_step'])"
done
python profiles/timeline_step.py cfg1 bf16 > gpurun_out/r02o/timeline_cfg1.txt 2>&1; sed -n 3,3p gpurun_out/r02o/timeline_cfg1.txt
