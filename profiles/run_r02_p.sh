#!/bin/bash
mkdir -p gpurun_out/r02p
python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/r02p/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02p/pytest.log
for fd in 1 0; do
for rep in 1 2; do
SLNLP_RNN_FUSED_DROPOUT=$fd python bench.py --steps 100 --warmup 10 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('mask-dropout $fd: cfg1', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step'], 'e2e', round(d['e2e']['value']))"
done
done
python bench.py --steps 50 --warmup 10 --precision fp32 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('fp32', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step'])"
python profiles/timeline_step.py cfg1 bf16 > gpurun_out/r02p/timeline_cfg1.txt 2>&1; sed -n 3,3p gpurun_out/r02p/timeline_cfg1.txt
