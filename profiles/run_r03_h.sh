#!/bin/bash
mkdir -p gpurun_out/r03h
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r03h/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r03h/pytest.log | head -20
timeout 900 python bench.py --workload cfg4 --steps 4 --warmup 3 --legs none --no-cpu-baseline > gpurun_out/r03h/bench_cfg4.json 2> gpurun_out/r03h/bench_cfg4.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r03h/bench_cfg4.json').read().strip().splitlines()[-1])
print('cfg4', round(d['value']), d['unit'], round(d['ms_per_step'],2), 'ms', d.get('launches_per_step')); print(d.get('roofline'))"
SLNLP_PDL=0 timeout 600 python profiles/kernel_table_step.py cfg4 bf16 > gpurun_out/r03h/table_cfg4_nopdl.txt 2>&1; grep -v Warn gpurun_out/r03h/table_cfg4_nopdl.txt | head -14
