#!/bin/bash
O=gpurun_out/r03
mkdir -p $O
python -m pytest tests -m gpu -q --timeout=900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest_gpu.log
( time python bench.py --steps 20 --warmup 5 ) > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 (all sub-records) rc=$?"
python bench.py --workload cfg4 --steps 6 --warmup 3 --legs none --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?"
SLNLP_PDL=0 python profiles/kernel_table_step.py cfg4 bf16 > $O/warm_cfg4_bf16.txt 2>&1
python profiles/prof_step_pair.py 16 4096 512 3 > $O/step_pair_final.txt 2>&1; cat $O/step_pair_final.txt
CMD4="python profiles/prof_step_pair.py 16 4096 512 1"
ncu --set full --clock-control none --import-source on -k regex:lstm_step_fwd -s 8 -c 1 -f -o $O/prof_step_fwd $CMD4 > $O/ncu_step_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_step_bwd -s 8 -c 1 -f -o $O/prof_step_bwd $CMD4 > $O/ncu_step_bwd.log 2>&1
python - <<P
import json
for f in ("bench_cfg1", "bench_cfg4"):
    d = json.loads(open("$O/" + f + ".json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round((d.get("e2e") or {}).get("value") or 0), " cpu", (d.get("cpu_baseline") or {}).get("value"))
    for k in ("fp32_path", "infer", "dp", "grid"):
        if k in d:
            print("   ", k, {kk: d[k].get(kk) for kk in ("value", "ms_per_step", "error", "search_seconds")})
P
