#!/bin/bash
# One gpurun call that regenerates the round's evidence: GPU parity suite, smoke, bench lines
# (cfg1 bf16 with CPU baseline, cfg1 fp32, cfg2, cfg3), the reference arm, an ncu launch list of the
# headline bench and full ncu captures of the three dominant kernels.  Outputs under gpurun_out/.
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_cfg1_bf16.json 2> gpurun_out/bench_cfg1_bf16.err; echo "bench cfg1 bf16 rc=$?"
python bench.py --precision fp32 --no-cpu-baseline > gpurun_out/bench_cfg1_fp32.json 2> gpurun_out/bench_cfg1_fp32.err; echo "bench cfg1 fp32 rc=$?"
python bench.py --workload cfg2 --steps 50 --warmup 5 > gpurun_out/bench_cfg2_bf16.json 2> gpurun_out/bench_cfg2_bf16.err; echo "bench cfg2 rc=$?"
python bench.py --workload cfg3 --steps 50 --warmup 5 > gpurun_out/bench_cfg3_bf16.json 2> gpurun_out/bench_cfg3_bf16.err; echo "bench cfg3 rc=$?"
python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/launches_cfg1_bf16.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu launches rc=$?"
CMD2="python profiles/prof_rnn_layer.py bf16 lstm"
$CMD2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rnn_persistent -s 2 -c 2 -f -o gpurun_out/prof_persist $CMD2 > gpurun_out/ncu_persist.log 2>&1
echo "ncu persistent rc=$?"
CMD3="python profiles/prof_gemm_one.py 0 1 3200 1024 128 tf32"
$CMD3 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tma -s 3 -c 1 -f -o gpurun_out/prof_gemm $CMD3 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
cat gpurun_out/plain2.log | tail -1
