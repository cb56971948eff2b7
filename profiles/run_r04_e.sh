#!/bin/bash
O=gpurun_out/r04e
mkdir -p $O
B="python bench.py --steps 30 --warmup 5 --legs none --no-cpu-baseline"
for i in 1 2; do
$B > $O/bench_fwdhead_$i.json 2> $O/err.txt
SLNLP_DEC_HEAD_BWD=1 $B > $O/bench_bothhead_$i.json 2> $O/err.txt
SLNLP_DEC_HEAD=0 $B > $O/bench_cell_$i.json 2> $O/err.txt
done
python - <<Q
import json
for f in ("bench_fwdhead_1", "bench_bothhead_1", "bench_cell_1", "bench_fwdhead_2", "bench_bothhead_2", "bench_cell_2"):
    d = json.loads(open("$O/" + f + ".json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]), "launches/step", d.get("launches_per_step"))
Q
