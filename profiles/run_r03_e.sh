#!/bin/bash
mkdir -p gpurun_out/r03e
SLNLP_TEST_TOL_SCALE=0.5 timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -k "fp32 or golden or oracle or estimator or fit_loop or factored or tf32x3 or cfg4" > gpurun_out/r03e/pytest_tight.log 2>&1; echo "tight pytest rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r03e/pytest_tight.log | head -20
for tc in 1 0; do
SLNLP_F32_TC=$tc python bench.py --steps 50 --warmup 10 --precision fp32 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('F32_TC=$tc fp32 cfg1', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step'])"
done
SLNLP_F32_TC=1 python bench.py --workload cfg2 --steps 30 --warmup 5 --precision fp32 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('fp32 cfg2', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms')"
