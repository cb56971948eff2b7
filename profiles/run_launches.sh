#!/bin/bash
# ncu launch list (per-launch device time) of the bench for one workload: bash profiles/run_launches.sh cfg1
WL=${1:-cfg1}
CMD="python bench.py --workload $WL --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$WL.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${WL}_bf16.csv $CMD > gpurun_out/ncu_$WL.log 2>&1
echo "ncu rc=$?"
python profiles/agg_launches.py gpurun_out/launches_${WL}_bf16.csv | head -45
