#!/bin/bash
mkdir -p gpurun_out/r03r
for k in 2 4 6 8; do
python bench.py --workload cfg5 --grid-fraction 0.5 --fits-per-gpu $k > gpurun_out/r03r/grid_slice_k$k.json 2> gpurun_out/r03r/grid_slice_k$k.err
python -c "
import json;d=json.loads(open('gpurun_out/r03r/grid_slice_k$k.json').read().strip().splitlines()[-1])
print('k=$k grid slice', round(d.get('value')), 'fits/hour', round(d.get('search_seconds') or (d.get('grid') or {}).get('search_seconds') or 0, 2), 's')"
done
