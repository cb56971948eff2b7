#!/bin/bash
mkdir -p gpurun_out/r03i
timeout 1500 python -m pytest tests/test_gpu_rnn_parity.py tests/test_gpu_kernels.py -m gpu -q --timeout=900 -k "large_batch or cfg4 or dropout_bf16" > gpurun_out/r03i/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r03i/pytest.log | head -20
timeout 900 python bench.py --workload cfg4 --steps 4 --warmup 3 --legs none --no-cpu-baseline > gpurun_out/r03i/bench_cfg4.json 2> gpurun_out/r03i/bench_cfg4.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r03i/bench_cfg4.json').read().strip().splitlines()[-1])
print('cfg4', round(d['value']), d['unit'], round(d['ms_per_step'],2), 'ms', d.get('launches_per_step'))"
SLNLP_PDL=0 timeout 600 python profiles/kernel_table_step.py cfg4 bf16 > gpurun_out/r03i/table_cfg4_nopdl.txt 2>&1; grep -v Warn gpurun_out/r03i/table_cfg4_nopdl.txt | head -8
