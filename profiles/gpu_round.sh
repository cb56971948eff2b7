#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (fp32 + bf16), ncu launch list and a
# full capture of the persistent recurrent kernels.  Outputs under gpurun_out/.
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py --precision fp32 > gpurun_out/bench_cfg1_fp32.json 2> gpurun_out/bench_cfg1_fp32.err; echo "bench fp32 rc=$?"
python bench.py --precision bf16 > gpurun_out/bench_cfg1_bf16.json 2> gpurun_out/bench_cfg1_bf16.err; echo "bench bf16 rc=$?"
CMD="python bench.py --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_cfg1_bf16.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu launches rc=$?"
CMD2="python profiles/prof_rnn_layer.py bf16 lstm"
$CMD2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rnn_persistent -s 2 -c 2 -f -o gpurun_out/prof_persist $CMD2 > gpurun_out/ncu_persist.log 2>&1
echo "ncu full rc=$?"
cat gpurun_out/plain2.log
