#!/bin/bash
mkdir -p gpurun_out/r03m
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_rnn_parity.py -m gpu -q --timeout=800 -k "cfg4 or large_batch" 2>&1 | tail -2
for B in 4096 2048 1024 512; do timeout 300 python profiles/prof_step_pair.py 16 $B 512 3; done
for S in 128 64 32; do echo "SLNLP_PAIR_SEQ=$S"; SLNLP_PAIR_SEQ=$S timeout 300 python profiles/prof_step_pair.py 16 1024 512 3; done
