#!/bin/bash
# One gpurun call that regenerates the evidence of the second half of round 2 (the large-batch family: CTA-pair GEMM,
# CTA-pair recurrent step kernels, split-operand fp32 GEMM): GPU parity suite, smoke, the default bench line with every
# sub-record (fp32_path, infer, dp = cfg4 at batch 4096, grid = the 810-fit cfg5 grid), cfg2 / cfg3 / cfg4 lines, the
# reference arm, warm per-kernel step tables (cfg1 both precisions, cfg4), GEMM / step-kernel microbenchmarks, the ncu
# launch list of the headline step and full ncu captures of the dominant kernels.  Outputs: gpurun_out/r03/.
set -o pipefail
O=gpurun_out/r03
mkdir -p $O
python -m pytest tests -m gpu -q --timeout=900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest_gpu.log
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
( time python bench.py --steps 20 --warmup 5 ) > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 (all sub-records) rc=$?"
python bench.py --workload cfg2 --steps 50 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?"
python bench.py --workload cfg3 --steps 50 --warmup 5 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench cfg3 rc=$?"
python bench.py --workload cfg4 --steps 6 --warmup 3 --legs none --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?"
SLNLP_PDL=0 python profiles/torch_prof_step.py cfg1 bf16 > $O/warm_cfg1_bf16.txt 2>&1
SLNLP_PDL=0 python profiles/torch_prof_step.py cfg1 fp32 > $O/warm_cfg1_fp32.txt 2>&1
SLNLP_PDL=0 python profiles/kernel_table_step.py cfg4 bf16 > $O/warm_cfg4_bf16.txt 2>&1
python profiles/timeline_step.py cfg1 bf16 > $O/timeline_cfg1.txt 2>&1
python profiles/bench_gemm_pair.py > $O/gemm_pair.txt 2>&1; echo "gemm pair rc=$?"
python profiles/prof_step_pair.py 16 4096 512 3 > $O/step_pair.txt 2>&1; SLNLP_PAIR_STEP=0 python profiles/prof_step_pair.py 16 4096 512 3 >> $O/step_pair.txt 2>&1
python profiles/bench_hbm_kernels.py > $O/hbm_kernels.txt 2>&1; echo "hbm kernels rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --legs none --no-cpu-baseline"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file $O/launches_cfg1_bf16.csv $CMD > $O/ncu.log 2>&1
echo "ncu launches rc=$?"
CMD2="python profiles/prof_rnn_layer.py bf16 lstm"
$CMD2 > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rnn_persistent -s 2 -c 2 -f -o $O/prof_persist $CMD2 > $O/ncu_persist.log 2>&1
echo "ncu persistent rc=$?"
CMD3="python profiles/bench_gemm_pair.py inproj"
ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 2 -c 1 -f -o $O/prof_gemm_pair $CMD3 > $O/ncu_gemm_pair.log 2>&1
echo "ncu gemm_pair rc=$?"
CMD4="python profiles/prof_step_pair.py 16 4096 512 1"
ncu --set full --clock-control none --import-source on -k regex:lstm_step_fwd -s 8 -c 1 -f -o $O/prof_step_fwd $CMD4 > $O/ncu_step_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_step_bwd -s 8 -c 1 -f -o $O/prof_step_bwd $CMD4 > $O/ncu_step_bwd.log 2>&1
echo "ncu step rc=$?"
python - <<P
import json
for f in ("bench_cfg1", "bench_cfg2", "bench_cfg3", "bench_cfg4", "bench_reference"):
    try:
        d = json.loads(open("$O/" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round((d.get("e2e") or {}).get("value") or 0), " cpu", (d.get("cpu_baseline") or {}).get("value"))
        for k in ("fp32_path", "infer", "dp", "grid"):
            if k in d:
                print("   ", k, {kk: d[k].get(kk) for kk in ("value", "ms_per_step", "error", "search_seconds")})
    except Exception as e:
        print(f, "ERR", e)
P
