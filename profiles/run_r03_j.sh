#!/bin/bash
mkdir -p gpurun_out/r03j
export SLNLP_PDL=0
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=1100 -k "gemm_bf16_cta_pair and (300 or 1000 or 8200) or tf32x3 and 262 or dropout_bf16 or concat_dirs" > gpurun_out/r03j/memcheck_kernels.log 2>&1; echo "memcheck kernels rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/r03j/memcheck_kernels.log | head
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rnn_parity.py -m gpu -q -x --timeout=1400 -k "large_batch_bf16_operand and 512" > gpurun_out/r03j/memcheck_module.log 2>&1; echo "memcheck module rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/r03j/memcheck_module.log | head
