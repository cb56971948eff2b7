#!/bin/bash
mkdir -p gpurun_out/r02y
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_baseline_golden.py tests/test_gpu_rnn_parity.py -m gpu -q -x --timeout=600 -k "cfg4 or large_batch" > gpurun_out/r02y/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02y/pytest.log
timeout 300 python profiles/prof_step_pair.py 16 4096 512 3; echo rc=$?
timeout 300 python profiles/prof_step_pair.py 16 512 512 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_step -s 10 -c 2 -o gpurun_out/r02y/step_fwd -f python profiles/prof_step_pair.py 16 4096 512 1 > gpurun_out/r02y/ncu1.log 2>&1; echo ncu rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_step -s 26 -c 2 -o gpurun_out/r02y/step_bwd -f python profiles/prof_step_pair.py 16 4096 512 1 > gpurun_out/r02y/ncu2.log 2>&1; echo ncu rc=$?
