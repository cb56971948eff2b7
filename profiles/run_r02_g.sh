#!/bin/bash
mkdir -p gpurun_out/r02g
for ps in 1 2; do
  SLNLP_PERSIST_PSEQ=$ps python -m pytest tests/test_gpu_kernels.py tests/test_gpu_rnn_parity.py -m gpu -q -x --timeout=900 -k "rnn or tensor_core or fused" > gpurun_out/r02g/pytest_ps$ps.log 2>&1
  echo "pytest pseq=$ps rc=$?"; tail -2 gpurun_out/r02g/pytest_ps$ps.log
done
for ps in 1 2 4; do
  SLNLP_PERSIST_PSEQ=$ps python profiles/prof_persist_phases.py lstm > gpurun_out/r02g/phases_lstm_ps$ps.txt 2>&1; echo "phases rc=$?"; cat gpurun_out/r02g/phases_lstm_ps$ps.txt
done
SLNLP_PERSIST_PSEQ=1 python profiles/prof_persist_phases.py gru > gpurun_out/r02g/phases_gru_ps1.txt 2>&1; grep variant gpurun_out/r02g/phases_gru_ps1.txt
for ps in 1 2 4; do
  SLNLP_PERSIST_PSEQ=$ps python bench.py --steps 50 --warmup 10 --legs none --no-cpu-baseline > gpurun_out/r02g/bench_ps$ps.json 2> gpurun_out/r02g/bench_ps$ps.err; echo "bench pseq=$ps rc=$?"
  python -c "import json;d=json.loads(open('gpurun_out/r02g/bench_ps$ps.json').read().strip().splitlines()[-1]);print('pseq $ps', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; per-timestep us', round(d['roofline']['us_per_timestep'],3))"
done
SLNLP_PERSIST_PSEQ=1 SLNLP_OVERLAP_DW=0 python bench.py --steps 50 --warmup 10 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('pseq 1 no-dw-overlap', round(d['value']), 'seq/s', round(d['ms_per_step'],4))"
