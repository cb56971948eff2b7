#!/bin/bash
O=gpurun_out/r07
mkdir -p $O
python -m pytest tests/test_gpu_transformer.py tests/test_gpu_estimator.py tests/test_gpu_baseline_golden.py -m gpu -q --timeout=900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest_gpu.log
python bench.py --workload cfg3 --steps 50 --warmup 5 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench cfg3 rc=$?"
SLNLP_PDL=0 python profiles/kernel_table_step.py cfg3 bf16 > $O/warm_cfg3_bf16.txt 2>&1
python - <<Q
import json
d = json.loads(open("$O/bench_cfg3.json").read().strip().splitlines()[-1])
print("cfg3", round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round((d.get("e2e") or {}).get("value") or 0), " cpu", (d.get("cpu_baseline") or {}).get("value"), "launches/step", d.get("launches_per_step"))
Q
