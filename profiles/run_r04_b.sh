#!/bin/bash
O=gpurun_out/r04b
mkdir -p $O
for i in 1 2; do
SLNLP_DEC_HEAD=0 SLNLP_DEC_CELL_BWD=0 python bench.py --steps 30 --warmup 5 --legs none --no-cpu-baseline > $O/bench_off_$i.json 2> $O/bench_off.err; echo "off rc=$?"
python bench.py --steps 30 --warmup 5 --legs none --no-cpu-baseline > $O/bench_on_$i.json 2> $O/bench_on.err; echo "on rc=$?"
done
SLNLP_DEC_HEAD=1 SLNLP_DEC_CELL_BWD=0 python bench.py --steps 30 --warmup 5 --legs none --no-cpu-baseline > $O/bench_head.json 2> $O/bench_head.err
SLNLP_DEC_HEAD=0 SLNLP_DEC_CELL_BWD=1 python bench.py --steps 30 --warmup 5 --legs none --no-cpu-baseline > $O/bench_cell.json 2> $O/bench_cell.err
python profiles/timeline_step.py cfg1 bf16 > $O/timeline_cfg1.txt 2>&1
SLNLP_PDL=0 python profiles/kernel_table_step.py cfg1 bf16 > $O/warm_cfg1_bf16.txt 2>&1
python - <<Q
import json
for f in ("bench_off_1", "bench_on_1", "bench_off_2", "bench_on_2", "bench_head", "bench_cell"):
    d = json.loads(open("$O/" + f + ".json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]), "launches/step", d.get("launches_per_step"))
Q
