#!/bin/bash
O=gpurun_out/r04a
mkdir -p $O
python -m pytest tests/test_gpu_rnn_parity.py tests/test_gpu_baseline_golden.py tests/test_gpu_estimator.py -m gpu -q --timeout=600 -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
python bench.py --steps 20 --warmup 5 --legs none --no-cpu-baseline > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench rc=$?"
python - <<Q
import json
d = json.loads(open("$O/bench_cfg1.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]), "launches/step", d.get("launches_per_step"))
Q
