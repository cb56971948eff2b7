#!/bin/bash
O=gpurun_out/r04h
mkdir -p $O
ncu --set full --import-source on --clock-control none -k "regex:dec_head|dec_cell_bwd" -s 9 -c 3 -f -o $O/dec_fused python bench.py --steps 4 --warmup 3 --legs none --no-cpu-baseline > $O/ncu.log 2>&1
echo "ncu rc=$?"; grep -c PROF $O/ncu.log
