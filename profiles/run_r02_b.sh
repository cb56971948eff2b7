#!/bin/bash
# round 2: default bench line with every sub-record; cfg5 slice at fits_per_gpu = 1, 2, 4, 6 (bf16) and 1 vs 4 (fp32)
mkdir -p gpurun_out/r02b
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02b/bench.json 2> gpurun_out/r02b/bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02b/bench.json; grep -v Warn gpurun_out/r02b/bench.err | tail -5
for k in 1 2 4 6; do
  ( time python bench.py --workload cfg5 --grid-fraction 0.1 --fits-per-gpu $k ) > gpurun_out/r02b/grid_k$k.json 2> gpurun_out/r02b/grid_k$k.err; echo "grid k=$k rc=$?"
done
for k in 1 4; do
  ( time python bench.py --workload cfg5 --grid-fraction 0.1 --fits-per-gpu $k --precision fp32 ) > gpurun_out/r02b/grid_fp32_k$k.json 2> gpurun_out/r02b/grid_fp32_k$k.err; echo "grid fp32 k=$k rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b/grid*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), 'fits/h', round(d['search_seconds'],1),'s')
    except Exception as e: print(f, 'ERR', e)
P
