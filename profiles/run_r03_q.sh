#!/bin/bash
mkdir -p gpurun_out/r03q
python profiles/prof_grid_fit.py 512 256 4 2>&1 | grep -E "ms per fit|torch.empty"
python profiles/prof_grid_fit.py 128 128 2 2>&1 | grep -E "ms per fit|torch.empty"
python bench.py --workload cfg5 --grid-fraction 0.5 > gpurun_out/r03q/grid_slice.json 2> gpurun_out/r03q/grid_slice.err; echo rc=$?
python -c "
import json;d=json.loads(open('gpurun_out/r03q/grid_slice.json').read().strip().splitlines()[-1]); g=d.get('grid') or d
print('grid slice', {k: g.get(k) for k in ('value','fits','search_seconds','fits_per_gpu')}, d.get('metric'), d.get('value'))"
timeout 600 python -m pytest tests/test_gpu_estimator.py -m gpu -q --timeout=500 2>&1 | tail -2
