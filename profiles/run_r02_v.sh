#!/bin/bash
mkdir -p gpurun_out/r02v
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_baseline_golden.py tests/test_gpu_rnn_parity.py -m gpu -q -x --timeout=600 -k "cfg4 or large_batch or dropout_bf16 or concat_dirs or gemm_bf16" > gpurun_out/r02v/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02v/pytest.log
timeout 600 python profiles/kernel_table_step.py cfg4 bf16 > gpurun_out/r02v/table_cfg4.txt 2>&1; echo rc=$?; grep -v Warn gpurun_out/r02v/table_cfg4.txt | head -24
