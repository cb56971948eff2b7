#!/bin/bash
O=gpurun_out/r05b
mkdir -p $O
B="python bench.py --workload cfg2 --steps 50 --warmup 5 --no-cpu-baseline --legs none"
for i in 1 2; do
SLNLP_DEC_HEAD=0 SLNLP_DEC_CELL_BWD=0 $B > $O/cfg2_off_$i.json 2> $O/err.txt
$B > $O/cfg2_on_$i.json 2> $O/err.txt
done
SLNLP_DEC_HEAD=0 $B > $O/cfg2_cell.json 2> $O/err.txt
SLNLP_DEC_CELL_BWD=0 $B > $O/cfg2_head.json 2> $O/err.txt
SLNLP_PDL=0 python profiles/kernel_table_step.py cfg2 bf16 > $O/warm_cfg2_bf16.txt 2>&1
python - <<Q
import json
for c in ("off_1", "on_1", "off_2", "on_2", "cell", "head"):
    d = json.loads(open("$O/cfg2_%s.json" % c).read().strip().splitlines()[-1])
    print(c, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  launches", d.get("launches_per_step"))
Q
grep "dec_head\|dec_cell" $O/warm_cfg2_bf16.txt
