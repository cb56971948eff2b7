#!/bin/bash
O=gpurun_out/r09
mkdir -p $O
python -m pytest tests/test_gpu_estimator.py -m gpu -q --timeout=300 -k "packed" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest.log
python bench.py --steps 20 --warmup 5 --legs none --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],4), d['launches_per_step'])"
