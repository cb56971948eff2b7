#!/bin/bash
# One gpurun call that regenerates the headline evidence after the fused decoder-side kernels (dec_head.cu, dec_cell.cu
# backward): GPU parity suite, smoke, the default bench line with every sub-record, the reference arm, the cfg2 (GRU) line,
# warm per-kernel step tables, the step timeline and the ncu launch list of the headline step.  Outputs: gpurun_out/r05/.
set -o pipefail
O=gpurun_out/r05
mkdir -p $O
python -m pytest tests -m gpu -q --timeout=900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest_gpu.log
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
( time python bench.py --steps 20 --warmup 5 ) > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 (all sub-records) rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?"
python bench.py --workload cfg2 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?"
SLNLP_PDL=0 python profiles/kernel_table_step.py cfg1 bf16 > $O/warm_cfg1_bf16.txt 2>&1
SLNLP_PDL=0 python profiles/kernel_table_step.py cfg1 fp32 > $O/warm_cfg1_fp32.txt 2>&1
python profiles/timeline_step.py cfg1 bf16 > $O/timeline_cfg1.txt 2>&1
CMD="python bench.py --steps 2 --warmup 3 --legs none --no-cpu-baseline"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 600 --csv --log-file $O/launches_cfg1_bf16.csv $CMD > $O/ncu.log 2>&1
echo "ncu launches rc=$?"
python - <<Q
import json
for f in ("bench_cfg1", "bench_cfg2", "bench_reference"):
    try:
        d = json.loads(open("$O/" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round((d.get("e2e") or {}).get("value") or 0), " cpu", (d.get("cpu_baseline") or {}).get("value"), "launches/step", d.get("launches_per_step"))
        for k in ("fp32_path", "infer", "dp", "grid"):
            if k in d:
                print("   ", k, {kk: d[k].get(kk) for kk in ("value", "ms_per_step", "ms_per_batch", "error", "search_seconds")})
    except Exception as e:
        print(f, "ERR", e)
Q
