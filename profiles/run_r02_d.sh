#!/bin/bash
# round 2: warp-specialised persistent kernel (parity for each variant, per-phase table), packed grid sweep
mkdir -p gpurun_out/r02d
for v in 0 3; do
  SLNLP_PERSIST_VAR=$v python -m pytest tests/test_gpu_kernels.py tests/test_gpu_rnn_parity.py tests/test_gpu_baseline_golden.py -m gpu -q -x --timeout=900 -k "rnn or tensor_core or cfg1 or cfg2 or fused" > gpurun_out/r02d/pytest_var$v.log 2>&1
  echo "pytest var=$v rc=$?"; tail -2 gpurun_out/r02d/pytest_var$v.log
done
python profiles/prof_persist_phases.py lstm > gpurun_out/r02d/phases_lstm.txt 2>&1; echo "phases rc=$?"; cat gpurun_out/r02d/phases_lstm.txt
python profiles/prof_persist_phases.py gru > gpurun_out/r02d/phases_gru.txt 2>&1; grep variant gpurun_out/r02d/phases_gru.txt
for v in 0 3; do
  SLNLP_PERSIST_VAR=$v python bench.py --steps 50 --warmup 10 --legs none --no-cpu-baseline > gpurun_out/r02d/bench_var$v.json 2> gpurun_out/r02d/bench_var$v.err; echo "bench var=$v rc=$?"
  python -c "import json;d=json.loads(open('gpurun_out/r02d/bench_var$v.json').read().strip().splitlines()[-1]);print('var $v', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; per-timestep us', round(d['roofline']['us_per_timestep'],3))"
done
for k in 1 2 4 6; do
  ( time python bench.py --workload cfg5 --grid-fraction 0.1 --fits-per-gpu $k ) > gpurun_out/r02d/grid_k$k.json 2> gpurun_out/r02d/grid_k$k.err; echo "grid k=$k rc=$?"
done
( time python bench.py --workload cfg5 --grid-fraction 0.1 --fits-per-gpu 1 --precision fp32 ) > gpurun_out/r02d/grid_fp32_k1.json 2> gpurun_out/r02d/grid_fp32_k1.err; echo "grid fp32 k=1 rc=$?"
( time python bench.py --workload cfg5 --grid-fraction 0.1 --fits-per-gpu 4 --precision fp32 ) > gpurun_out/r02d/grid_fp32_k4.json 2> gpurun_out/r02d/grid_fp32_k4.err; echo "grid fp32 k=4 rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d/grid*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), 'fits/h', round(d['search_seconds'],1),'s')
    except Exception as e: print(f, 'ERR', e)
P
grep -v "Warn\|warn" gpurun_out/r02d/grid_k4.err | grep -v "^frame\|^$" | tail -12
