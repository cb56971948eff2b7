#!/bin/bash
O=gpurun_out/r04d
mkdir -p $O
python -m pytest tests/test_gpu_rnn_parity.py tests/test_gpu_baseline_golden.py -m gpu -q --timeout=600 -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
B="python bench.py --steps 30 --warmup 5 --legs none --no-cpu-baseline"
for i in 1 2; do
SLNLP_DEC_HEAD=0 SLNLP_DEC_CELL_BWD=0 $B > $O/bench_off_$i.json 2> $O/err.txt
$B > $O/bench_on_$i.json 2> $O/err.txt
done
SLNLP_DEC_HEAD=0 $B > $O/bench_cell.json 2> $O/err.txt
SLNLP_DEC_CELL_BWD=0 $B > $O/bench_head.json 2> $O/err.txt
SLNLP_DEC_HEAD_FUSE=1 $B > $O/bench_fuse.json 2> $O/err.txt
python profiles/timeline_step.py cfg1 bf16 > $O/timeline_cfg1.txt 2>&1
SLNLP_PDL=0 python profiles/kernel_table_step.py cfg1 bf16 > $O/warm_cfg1_bf16.txt 2>&1
python - <<Q
import json
for f in ("bench_off_1", "bench_on_1", "bench_off_2", "bench_on_2", "bench_cell", "bench_head", "bench_fuse"):
    d = json.loads(open("$O/" + f + ".json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]), "launches/step", d.get("launches_per_step"))
Q
grep "dec_head\|dec_cell" $O/warm_cfg1_bf16.txt
