#!/bin/bash
mkdir -p gpurun_out/r03d
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r03d/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r03d/pytest.log | head -20
SLNLP_TEST_TOL_SCALE=0.5 timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -k "fp32 or golden or oracle or estimator or fit_loop or factored" > gpurun_out/r03d/pytest_tight.log 2>&1; echo "tight pytest rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r03d/pytest_tight.log | head -20
SLNLP_F32_TC=0 SLNLP_TEST_TOL_SCALE=0.5 timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -k "fp32 or golden or oracle or estimator or fit_loop or factored" > gpurun_out/r03d/pytest_tight_fma.log 2>&1; echo "tight (fp32-FMA GEMM) pytest rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r03d/pytest_tight_fma.log | head -20
