#!/bin/bash
mkdir -p gpurun_out/r03s
SLNLP_PDL=0 timeout 600 python profiles/kernel_table_step.py cfg3 bf16 > gpurun_out/r03s/table_cfg3.txt 2>&1; grep -v Warn gpurun_out/r03s/table_cfg3.txt | head -32
SLNLP_PDL=0 timeout 600 python profiles/kernel_table_step.py cfg2 bf16 > gpurun_out/r03s/table_cfg2.txt 2>&1; grep -v Warn gpurun_out/r03s/table_cfg2.txt | head -24
