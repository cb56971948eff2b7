#!/bin/bash
mkdir -p gpurun_out/r03l
for i in 1 2; do
python bench.py --steps 100 --warmup 10 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cfg1', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step'], 'e2e', round(d['e2e']['value']))"
done
python profiles/timeline_step.py cfg1 bf16 > gpurun_out/r03l/timeline_cfg1.txt 2>&1; sed -n 3,3p gpurun_out/r03l/timeline_cfg1.txt
for wl in cfg2 cfg3; do
python bench.py --workload $wl --steps 50 --warmup 10 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$wl', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms')"
done
python bench.py --steps 50 --warmup 10 --precision fp32 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('fp32 cfg1', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms')"
timeout 900 python -m pytest tests/test_gpu_rnn_parity.py tests/test_gpu_baseline_golden.py tests/test_gpu_estimator.py tests/test_gpu_transformer.py -m gpu -q --timeout=800 2>&1 | tail -2
