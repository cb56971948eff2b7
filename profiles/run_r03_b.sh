#!/bin/bash
mkdir -p gpurun_out/r03b
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/r03b/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r03b/pytest.log
for tc in 1 0; do
SLNLP_F32_TC=$tc python bench.py --steps 50 --warmup 10 --precision fp32 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('F32_TC=$tc fp32 cfg1', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step'])"
done
for wl in cfg2 cfg3; do
SLNLP_F32_TC=1 python bench.py --workload $wl --steps 30 --warmup 5 --precision fp32 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('F32_TC=1 fp32 $wl', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms')"
done
