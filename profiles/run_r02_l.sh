#!/bin/bash
mkdir -p gpurun_out/r02l
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  for rep in 1 2; do
  SLNLP_RNN_EXTRAS=$1 SLNLP_RNN_FUSED_DROPOUT=$2 python bench.py --steps 100 --warmup 10 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('extras $1 fused-dropout $2: cfg1', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step'], 'rnn us/step', round(d['roofline']['us_per_timestep'],3))"
  done
done
SLNLP_RNN_EXTRAS=1 SLNLP_RNN_FUSED_DROPOUT=0 python profiles/timeline_step.py cfg1 bf16 > gpurun_out/r02l/timeline_cfg1.txt 2>&1; sed -n 3,3p gpurun_out/r02l/timeline_cfg1.txt
