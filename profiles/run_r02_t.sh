#!/bin/bash
mkdir -p gpurun_out/r02t
timeout 600 python profiles/kernel_table_step.py cfg4 bf16 > gpurun_out/r02t/table_cfg4.txt 2>&1; echo rc=$?; grep -v Warn gpurun_out/r02t/table_cfg4.txt | head -40
