#!/bin/bash
# last run of round 2 after the size threshold of the fused cell backward: parity of the RNN modules + estimator, the
# default bench line (all sub-records), the cfg2 line
O=gpurun_out/r06
mkdir -p $O
python -m pytest tests/test_gpu_rnn_parity.py tests/test_gpu_estimator.py tests/test_gpu_baseline_golden.py -m gpu -q --timeout=900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest_gpu.log
( time python bench.py --steps 20 --warmup 5 ) > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 (all sub-records) rc=$?"
python bench.py --workload cfg2 --steps 50 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?"
python - <<Q
import json
for f in ("bench_cfg1", "bench_cfg2"):
    try:
        d = json.loads(open("$O/" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round((d.get("e2e") or {}).get("value") or 0), " cpu", (d.get("cpu_baseline") or {}).get("value"), "launches/step", d.get("launches_per_step"))
        for k in ("fp32_path", "infer", "dp", "grid"):
            if k in d:
                print("   ", k, {kk: d[k].get(kk) for kk in ("value", "ms_per_step", "ms_per_batch", "error", "search_seconds")})
    except Exception as e:
        print(f, "ERR", e)
Q
