#!/bin/bash
mkdir -p gpurun_out/r02w
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_baseline_golden.py tests/test_gpu_rnn_parity.py -m gpu -q -x --timeout=600 -k "cfg4 or large_batch or dropout_bf16 or concat_dirs" > gpurun_out/r02w/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02w/pytest.log
SLNLP_PDL=0 timeout 600 python profiles/kernel_table_step.py cfg4 bf16 > gpurun_out/r02w/table_cfg4_nopdl.txt 2>&1; echo rc=$?; grep -v Warn gpurun_out/r02w/table_cfg4_nopdl.txt | head -14
timeout 600 python profiles/kernel_table_step.py cfg4 bf16 > gpurun_out/r02w/table_cfg4.txt 2>&1; echo rc=$?; grep -v Warn gpurun_out/r02w/table_cfg4.txt | head -6
SLNLP_PAIR_STEP=0 SLNLP_PDL=0 timeout 600 python profiles/kernel_table_step.py cfg4 bf16 > gpurun_out/r02w/table_cfg4_tcstep_nopdl.txt 2>&1; echo rc=$?; grep -v Warn gpurun_out/r02w/table_cfg4_tcstep_nopdl.txt | head -6
