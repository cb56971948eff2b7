#!/bin/bash
mkdir -p gpurun_out/r02k
python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/r02k/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02k/pytest.log
for ex in 0 1; do
  SLNLP_RNN_EXTRAS=$ex python bench.py --steps 50 --warmup 10 --legs fp32_path --no-cpu-baseline > gpurun_out/r02k/bench_ex$ex.json 2> gpurun_out/r02k/bench_ex$ex.err; echo "bench extras=$ex rc=$?"
  python -c "import json;d=json.loads(open('gpurun_out/r02k/bench_ex$ex.json').read().strip().splitlines()[-1]);print('extras $ex: cfg1', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step']); f=d['fp32_path']; print('   fp32', round(f['value']), round(f['ms_per_step'],4), f['launches_per_step'])"
done
for wl in cfg2 cfg3; do
python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$wl', round(d['value']), 'seq/s', round(d['ms_per_step'],4), 'launches', d['launches_per_step'])"
done
python profiles/timeline_step.py cfg1 bf16 > gpurun_out/r02k/timeline_cfg1.txt 2>&1; head -3 gpurun_out/r02k/timeline_cfg1.txt | tail -1
