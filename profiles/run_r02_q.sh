#!/bin/bash
mkdir -p gpurun_out/r02q
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=300 -k "gemm_bf16 or cast_bf16" > gpurun_out/r02q/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02q/pytest.log
for eg in 1 2; do for dbg in 0 1 2; do
echo "== EG=$eg DBG=$dbg"
SLNLP_PAIR_EG=$eg SLNLP_PAIR_DBG=$dbg timeout 600 python profiles/bench_gemm_pair.py inproj dW_ih 2>&1 | cut -c1-150
done; done
