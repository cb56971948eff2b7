#!/bin/bash
mkdir -p gpurun_out/r02j
python -m pytest tests/test_gpu_estimator.py tests/test_gpu_rnn_parity.py -m gpu -q -x --timeout=900 > gpurun_out/r02j/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02j/pytest.log
python profiles/timeline_step.py cfg1 bf16 > gpurun_out/r02j/timeline_cfg1.txt 2>&1; echo "timeline rc=$?"; tail -5 gpurun_out/r02j/timeline_cfg1.txt
SLNLP_PDL=0 python profiles/timeline_step.py cfg1 bf16 > gpurun_out/r02j/timeline_cfg1_nopdl.txt 2>&1
SLNLP_PDL=0 python profiles/torch_prof_step.py cfg2 bf16 > gpurun_out/r02j/warm_cfg2.txt 2>&1; head -24 gpurun_out/r02j/warm_cfg2.txt
for k in 1 4; do
  python bench.py --workload cfg5 --grid-fraction 0.1 --fits-per-gpu $k --precision fp32 > gpurun_out/r02j/grid_fp32_k$k.json 2> gpurun_out/r02j/grid_fp32_k$k.err; echo "grid fp32 k=$k rc=$?"
done
python - <<'P'
import json
def load(f): return json.loads(open('gpurun_out/r02j/'+f).read().strip().splitlines()[-1])
a, b = load('grid_fp32_k1.json'), load('grid_fp32_k4.json')
print('fits/h', round(a['value']), round(b['value']), 'max |mean_test_score diff|', max(abs(x - y) for x, y in zip(a['mean_test_score'], b['mean_test_score'])))
P
