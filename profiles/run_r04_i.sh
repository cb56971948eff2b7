#!/bin/bash
O=gpurun_out/r04i
mkdir -p $O
python -m pytest tests/test_gpu_rnn_parity.py -m gpu -q --timeout=600 -x -k "fused or golden or baseline_size" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
B="python bench.py --steps 30 --warmup 5 --legs none --no-cpu-baseline"
for i in 1 2; do
SLNLP_DEC_HEAD=0 SLNLP_DEC_CELL_BWD=0 $B > $O/bench_off_$i.json 2> $O/err.txt
$B > $O/bench_on_$i.json 2> $O/err.txt
done
SLNLP_PDL=0 python profiles/kernel_table_step.py cfg1 bf16 > $O/warm_cfg1_bf16.txt 2>&1
python - <<Q
import json
for c in ("off_1", "on_1", "off_2", "on_2"):
    d = json.loads(open("$O/bench_%s.json" % c).read().strip().splitlines()[-1])
    print(c, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]))
Q
grep "dec_head\|dec_cell" $O/warm_cfg1_bf16.txt
