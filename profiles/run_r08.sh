#!/bin/bash
# cfg5 grid (810 fits) on one GPU: worker processes x threads per process
O=gpurun_out/r08
mkdir -p $O
python -m pytest tests/test_gpu_estimator.py -m gpu -q --timeout=600 -k "packed" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest.log
for pk in "4 2" "3 4" "6 1"; do
set -- $pk
python bench.py --workload cfg5 --procs-per-gpu $1 --fits-per-gpu $2 > $O/grid_p$1_k$2.json 2> $O/grid_p$1_k$2.err; echo "grid p$1 k$2 rc=$?"
done
python - <<Q
import json, glob
for f in sorted(glob.glob("$O/grid_p*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["value"]), d["unit"], round(d["search_seconds"], 1), "s  launches", d.get("gpu_launches"), "best", round(d["best_score"], 4))
    except Exception as e:
        print(f, "ERR", e)
Q
