#!/bin/bash
mkdir -p gpurun_out/r02h
python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/r02h/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02h/pytest.log
for f in 0 1; do
  SLNLP_DEC_FUSED=$f python bench.py --steps 50 --warmup 10 --legs none --no-cpu-baseline > gpurun_out/r02h/bench_fused$f.json 2> gpurun_out/r02h/bench_fused$f.err; echo "bench fused=$f rc=$?"
  python -c "import json;d=json.loads(open('gpurun_out/r02h/bench_fused$f.json').read().strip().splitlines()[-1]);print('dec fused $f', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step'])"
done
python bench.py --steps 30 --warmup 5 --legs fp32_path --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('fp32', d['fp32_path'])"
for wl in cfg2 cfg3; do
python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$wl', round(d['value']), 'seq/s', round(d['ms_per_step'],4), 'launches', d['launches_per_step'])"
done
SLNLP_PDL=0 python profiles/torch_prof_step.py cfg1 bf16 > gpurun_out/r02h/warm_cfg1.txt 2>&1; head -40 gpurun_out/r02h/warm_cfg1.txt
