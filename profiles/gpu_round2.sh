#!/bin/bash
# One gpurun call that regenerates round 2's evidence: GPU parity suite, smoke, the default bench line with
# every sub-record (fp32_path, infer, dp = cfg4 at batch 4096, grid = the full 810-fit cfg5 grid), cfg2 / cfg3
# lines, the reference arm, the per-phase cycle table of the persistent kernel, warm per-kernel step tables,
# an ncu launch list of the headline step and full ncu captures of the dominant kernels.  Outputs: gpurun_out/r02/.
set -o pipefail
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -m gpu -x -q --timeout=900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest_gpu.log
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
( time python bench.py --steps 20 --warmup 5 ) > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 (all sub-records) rc=$?"
python bench.py --workload cfg2 --steps 50 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?"
python bench.py --workload cfg3 --steps 50 --warmup 5 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench cfg3 rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?"
python profiles/prof_persist_phases.py lstm > $O/persist_phases_lstm.txt 2>&1; python profiles/prof_persist_phases.py gru > $O/persist_phases_gru.txt 2>&1
SLNLP_PDL=0 python profiles/torch_prof_step.py cfg1 bf16 > $O/warm_cfg1_bf16.txt 2>&1
SLNLP_PDL=0 python profiles/torch_prof_step.py cfg1 fp32 > $O/warm_cfg1_fp32.txt 2>&1
python profiles/bench_hbm_kernels.py > $O/hbm_kernels.txt 2>&1; echo "hbm kernels rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --legs none --no-cpu-baseline"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file $O/launches_cfg1_bf16.csv $CMD > $O/ncu.log 2>&1
echo "ncu launches rc=$?"
CMD2="python profiles/prof_rnn_layer.py bf16 lstm"
$CMD2 > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rnn_persistent -s 2 -c 2 -f -o $O/prof_persist $CMD2 > $O/ncu_persist.log 2>&1
echo "ncu persistent rc=$?"
CMD3="python profiles/prof_rnn_layer.py fp32 lstm"
$CMD3 > $O/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rnn_pf32 -s 2 -c 2 -f -o $O/prof_pf32 $CMD3 > $O/ncu_pf32.log 2>&1
echo "ncu pf32 rc=$?"
CMD4="python profiles/prof_gemm_one.py 0 1 3200 1024 128 tf32"
$CMD4 > $O/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tma -s 3 -c 1 -f -o $O/prof_gemm $CMD4 > $O/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
tail -1 $O/plain2.log; tail -1 $O/plain3.log
python - <<P
import json
for f in ("bench_cfg1", "bench_cfg2", "bench_cfg3", "bench_reference"):
    try:
        d = json.loads(open("$O/" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]), " cpu", (d.get("cpu_baseline") or {}).get("value"))
        for k in ("fp32_path", "infer", "dp", "grid"):
            if k in d:
                print("   ", k, {kk: d[k].get(kk) for kk in ("value", "ms_per_step", "error", "search_seconds")})
    except Exception as e:
        print(f, "ERR", e)
P
