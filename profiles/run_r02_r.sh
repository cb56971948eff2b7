#!/bin/bash
mkdir -p gpurun_out/r02r
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_baseline_golden.py -m gpu -q -x --timeout=600 -k "gemm_bf16 or cast_bf16 or cfg4" > gpurun_out/r02r/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02r/pytest.log
for pair in 1 0; do
SLNLP_PAIR=$pair timeout 900 python bench.py --workload cfg4 --steps 4 --warmup 3 --legs none --no-cpu-baseline > gpurun_out/r02r/bench_cfg4_pair$pair.json 2> gpurun_out/r02r/bench_cfg4_pair$pair.err; echo "bench pair=$pair rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r02r/bench_cfg4_pair$pair.json').read().strip().splitlines()[-1])
print('pair=$pair cfg4', round(d['value']), d['unit'], round(d['ms_per_step'],2), 'ms', d.get('launches_per_step'), d.get('roofline'))"
done
