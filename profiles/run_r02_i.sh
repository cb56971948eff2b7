#!/bin/bash
mkdir -p gpurun_out/r02i
python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/r02i/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02i/pytest.log
python bench.py --steps 50 --warmup 10 --legs fp32_path --no-cpu-baseline > gpurun_out/r02i/bench.json 2> gpurun_out/r02i/bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r02i/bench.json').read().strip().splitlines()[-1]);print('cfg1', round(d['value']), 'seq/s', round(d['ms_per_step'],4),'ms; launches', d['launches_per_step']); f=d['fp32_path']; print('fp32', round(f['value']), round(f['ms_per_step'],4), f['launches_per_step'])"
SLNLP_PERSIST_F32=0 python bench.py --steps 30 --warmup 5 --precision fp32 --legs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('fp32 per-step kernels', round(d['value']), 'seq/s', round(d['ms_per_step'],4), 'launches', d['launches_per_step'])"
for wl in cfg2 cfg3; do
python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$wl', round(d['value']), 'seq/s', round(d['ms_per_step'],4), 'launches', d['launches_per_step'])"
done
python profiles/prof_rnn_layer.py fp32 lstm 128 | tail -2
python profiles/prof_rnn_layer.py fp32 gru 128 | tail -1
SLNLP_PDL=0 python profiles/torch_prof_step.py cfg1 fp32 > gpurun_out/r02i/warm_cfg1_fp32.txt 2>&1; head -30 gpurun_out/r02i/warm_cfg1_fp32.txt
