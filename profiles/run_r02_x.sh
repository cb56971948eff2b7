#!/bin/bash
mkdir -p gpurun_out/r02x
timeout 300 python profiles/prof_step_pair.py 16 4096 512 3; echo rc=$?
SLNLP_PAIR_STEP=0 timeout 300 python profiles/prof_step_pair.py 16 4096 512 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_step -s 20 -c 4 -o gpurun_out/r02x/step_pair -f python profiles/prof_step_pair.py 16 4096 512 1 > gpurun_out/r02x/ncu.log 2>&1; echo ncu rc=$?; tail -3 gpurun_out/r02x/ncu.log
