#!/bin/bash
O=gpurun_out/r04f
mkdir -p $O
python profiles/bench_gemm_dw.py > $O/gemm_dw.txt 2>&1; cat $O/gemm_dw.txt
B="python bench.py --steps 30 --warmup 5 --legs none --no-cpu-baseline"
for cap in 0 1 2 4 8; do
if [ $cap = 0 ]; then $B > $O/bench_cap$cap.json 2> $O/err.txt; else SLNLP_SPLITK_MAX=$cap $B > $O/bench_cap$cap.json 2> $O/err.txt; fi
done
python - <<Q
import json
for c in (0, 1, 2, 4, 8):
    d = json.loads(open("$O/bench_cap%d.json" % c).read().strip().splitlines()[-1])
    print("cap", c, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]))
Q
