#!/bin/bash
# full ncu capture of the few-query attention kernels inside one cfg3 step (eager launches)
rm -f gpurun_out/*.ncu-rep
ncu --set full --import-source on --clock-control none -k regex:mha_small -s 6 -c 6 -o gpurun_out/mha_small python bench.py --workload cfg3 --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_small.log 2>&1
echo "ncu rc=$?"
