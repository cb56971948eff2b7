#!/bin/bash
# parity suite, bench lines with and without the small-dW side branch, one full ncu capture of layernorm_bwd
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
line() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2], round(d['value']),'seq/s', round(d['ms_per_step'],3),'ms', d['launches_per_step'],'launches e2e', round(d['e2e']['value']))" $1 "$2"; }
for wl in cfg1 cfg2 cfg3; do
python bench.py --workload $wl --precision bf16 --steps 30 --warmup 3 --no-cpu-baseline 2>gpurun_out/b_${wl}_bf16.err > gpurun_out/b_${wl}_bf16.json; line gpurun_out/b_${wl}_bf16.json "$wl"
done
for wl in cfg1 cfg2; do
SLNLP_OVERLAP_SMALL=0 python bench.py --workload $wl --precision bf16 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/b_${wl}_nosmall.json; line gpurun_out/b_${wl}_nosmall.json "$wl no-small-overlap"
done
python bench.py --workload cfg1 --precision fp32 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/b_cfg1_fp32.json; line gpurun_out/b_cfg1_fp32.json "cfg1 fp32"
rm -f gpurun_out/*.ncu-rep
ncu --set full --import-source on --clock-control none -k regex:layernorm_bwd -s 30 -c 1 -o gpurun_out/ln_bwd python bench.py --workload cfg3 --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ln.log 2>&1
echo "ncu rc=$?"
