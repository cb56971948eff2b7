#!/bin/bash
# RNN parity suite + cfg1/cfg2 with the decoder cell on the tensor-core step kernel (default) and on the fp32 one
python -m pytest tests/test_gpu_rnn_parity.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -2
line() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2], round(d['value']),'seq/s', round(d['ms_per_step'],3),'ms', d['launches_per_step'],'launches e2e', round(d['e2e']['value']))" $1 "$2"; }
for v in 1 0; do for wl in cfg1 cfg2; do
SLNLP_DEC_TC=$v python bench.py --workload $wl --precision bf16 --steps 30 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab3.err > gpurun_out/ab3.json; line gpurun_out/ab3.json "$wl dec_tc=$v"
done; done
SLNLP_PDL=0 python profiles/torch_prof_step.py cfg1 bf16 2>&1 | grep -A30 "us/step" | cut -c1-120
