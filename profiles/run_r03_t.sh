#!/bin/bash
for v in 0 1 0 1; do
SLNLP_TR_SIDE_BIG=$v python bench.py --workload cfg3 --steps 50 --warmup 10 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('side_big=$v cfg3', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms')"
done
SLNLP_TR_SIDE_BIG=1 timeout 600 python -m pytest tests/test_gpu_transformer.py -m gpu -q --timeout=500 2>&1 | tail -2
