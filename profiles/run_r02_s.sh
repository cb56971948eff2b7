#!/bin/bash
mkdir -p gpurun_out/r02s
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_baseline_golden.py -m gpu -q -x --timeout=600 -k "gemm_bf16 or cfg4" > gpurun_out/r02s/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02s/pytest.log
for bf in 1 0; do
SLNLP_BF_STEP=$bf timeout 900 python bench.py --workload cfg4 --steps 4 --warmup 3 --legs none --no-cpu-baseline > gpurun_out/r02s/bench_cfg4_bf$bf.json 2> gpurun_out/r02s/bench_cfg4_bf$bf.err; echo "bench bf=$bf rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r02s/bench_cfg4_bf$bf.json').read().strip().splitlines()[-1])
print('bf_step=$bf cfg4', round(d['value']), d['unit'], round(d['ms_per_step'],2), 'ms', d.get('launches_per_step'), d.get('roofline'))"
done
