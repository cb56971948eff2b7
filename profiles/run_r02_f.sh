#!/bin/bash
mkdir -p gpurun_out/r02f
for v in 8 9; do
  SLNLP_PERSIST_VAR=$v python -m pytest tests/test_gpu_kernels.py tests/test_gpu_rnn_parity.py -m gpu -q -x --timeout=900 -k "rnn or tensor_core or fused" > gpurun_out/r02f/pytest_var$v.log 2>&1
  echo "pytest var=$v rc=$?"; tail -2 gpurun_out/r02f/pytest_var$v.log
done
python profiles/prof_persist_phases.py lstm > gpurun_out/r02f/phases_lstm.txt 2>&1; echo "phases rc=$?"; cat gpurun_out/r02f/phases_lstm.txt
python profiles/prof_persist_phases.py gru > gpurun_out/r02f/phases_gru.txt 2>&1; grep variant gpurun_out/r02f/phases_gru.txt
for v in 0 1 8 9; do
  SLNLP_PERSIST_VAR=$v python bench.py --steps 50 --warmup 10 --legs none --no-cpu-baseline > gpurun_out/r02f/bench_var$v.json 2> gpurun_out/r02f/bench_var$v.err; echo "bench var=$v rc=$?"
  python -c "import json;d=json.loads(open('gpurun_out/r02f/bench_var$v.json').read().strip().splitlines()[-1]);print('var $v', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; per-timestep us', round(d['roofline']['us_per_timestep'],3))"
done
