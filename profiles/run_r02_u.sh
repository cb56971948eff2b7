#!/bin/bash
mkdir -p gpurun_out/r02u
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_baseline_golden.py -m gpu -q -x --timeout=600 -k "cfg4" > gpurun_out/r02u/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02u/pytest.log
timeout 600 python profiles/kernel_table_step.py cfg4 bf16 > gpurun_out/r02u/table_cfg4.txt 2>&1; echo rc=$?; grep -v Warn gpurun_out/r02u/table_cfg4.txt | head -16
