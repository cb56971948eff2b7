#!/bin/bash
mkdir -p gpurun_out/r02z
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_rnn_parity.py -m gpu -q -x --timeout=600 -k "cfg4 or large_batch" > gpurun_out/r02z/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z/pytest.log
timeout 300 python profiles/prof_step_pair.py 16 4096 512 3; echo rc=$?
SLNLP_PAIR_SEQ=128 timeout 300 python profiles/prof_step_pair.py 16 4096 512 3
timeout 300 python profiles/prof_step_pair.py 16 512 512 3
timeout 900 python bench.py --workload cfg4 --steps 4 --warmup 3 --legs none --no-cpu-baseline > gpurun_out/r02z/bench_cfg4.json 2> gpurun_out/r02z/bench_cfg4.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r02z/bench_cfg4.json').read().strip().splitlines()[-1])
print('cfg4', round(d['value']), d['unit'], round(d['ms_per_step'],2), 'ms', d.get('launches_per_step'))"
