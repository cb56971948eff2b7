#!/bin/bash
mkdir -p gpurun_out/r03c
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=600 -k "tf32x3 or gemm_tf32" > gpurun_out/r03c/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r03c/pytest.log | head -30
