#!/bin/bash
mkdir -p gpurun_out/r03g
SLNLP_PDL=0 timeout 600 python profiles/kernel_table_step.py cfg4 bf16 > gpurun_out/r03g/table_cfg4_nopdl.txt 2>&1; echo rc=$?; grep -v Warn gpurun_out/r03g/table_cfg4_nopdl.txt | head -44
