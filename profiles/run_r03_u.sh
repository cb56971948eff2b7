#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_rnn_parity.py tests/test_gpu_dp.py -m gpu -q --timeout=800 -k "cfg4 or large_batch or dp" 2>&1 | tail -3
for v in 1 0; do
SLNLP_GATES_BF16=$v timeout 900 python bench.py --workload cfg4 --steps 6 --warmup 3 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('gates_bf16=$v cfg4', round(d['value']), d['unit'], round(d['ms_per_step'],2), 'ms')"
done
